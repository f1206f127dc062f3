"""Probe: may a K-major SWIZZLE_128B UMMA descriptor start at a row offset that is not a multiple of 8?
The A tile (one image row of 128 pixels) is loaded `r` pixels early and the descriptor starts r rows in, so
the result must equal the normal conv except for the last r pixels of every 128-pixel tile."""
import os, subprocess, sys, json
if len(sys.argv) > 1:
    import torch, torch.nn.functional as F
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from pytorch_ddp_resnet_b200 import ops, _lib
    r = int(os.environ.get("B200_PROBE_ROWOFF", "0"))
    torch.manual_seed(0)
    N, H, W, C, K = 2, 4, 128, 64, 64
    x = torch.randn(N, H, W, C, device="cuda").bfloat16()
    w = (torch.randn(K, 3, 3, C, device="cuda") * 0.05).bfloat16()
    y = ops.conv_fprop(x, w, 1, 1, algo=_lib.ALGO_TC).float()
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.permute(0, 3, 1, 2).float(), padding=1).permute(0, 2, 3, 1)
    ok = slice(0, W - r)
    err = ((y[:, :, ok] - ref[:, :, ok]).norm() / ref[:, :, ok].norm()).item()
    print(json.dumps({"rowoff": r, "baseoff": int(os.environ.get("B200_PROBE_BASEOFF", "0")), "rel_l2": err}))
else:
    for r, b in [(0, 0), (8, 0), (1, 0), (1, 1), (2, 2), (3, 3), (5, 5), (7, 7), (9, 1), (17, 1)]:
        env = dict(os.environ, B200_CONV_PAIR="0", B200_CONV_CLUSTER="1", B200_PROBE_ROWOFF=str(r), B200_PROBE_BASEOFF=str(b))
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, timeout=120)
        print(out.stdout.strip() or out.stderr.strip()[-300:], flush=True)
