"""Warm CUDA-event timing of the HBM-bound kernels at the WRN-28-10 shapes, rotating over buffer sets
larger than L2, in a CUDA graph (no launch gaps). Prints achieved GB/s against algorithmic bytes."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_ddp_resnet_b200 import ops

torch.manual_seed(0)
dev = "cuda"
out = []
for (N, H, C) in [(128, 32, 160), (128, 16, 320), (128, 8, 640)]:
    tensor_mb = N * H * H * C * 2 / 1e6
    nset = max(2, int(400 / (3 * tensor_mb)) + 1)      # >= 400 MB touched per rotation
    xs = [torch.randn(N, H, H, C, device=dev).bfloat16() for _ in range(nset)]
    dys = [torch.randn(N, H, H, C, device=dev).bfloat16() for _ in range(nset)]
    gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
    mean, invstd = ops.bn_stats(xs[0], 1e-5)
    ays = [ops.bn_act_fwd(x, mean, invstd, gamma, beta, relu=True, dropout_p=0.3, seed=1) for x in xs]
    mks = [ops.bn_act_fwd(x, mean, invstd, gamma, beta, relu=True, dropout_p=0.3, seed=1, want_mask=True)[1] for x in xs]

    cases = {
        "bn_stats":            (1, lambda i: ops.bn_stats(xs[i], 1e-5)),
        "bn_act_fwd":          (2, lambda i: ops.bn_act_fwd(xs[i], mean, invstd, gamma, beta, relu=True)),
        "bn_act_fwd_dropout":  (2, lambda i: ops.bn_act_fwd(xs[i], mean, invstd, gamma, beta, relu=True, dropout_p=0.3, seed=1)),
        "bn_act_fwd_drop_mask": (2.0625, lambda i: ops.bn_act_fwd(xs[i], mean, invstd, gamma, beta, relu=True, dropout_p=0.3, seed=1, want_mask=True)),
        "bn_act_bwd_mask_drop": (5.125, lambda i: ops.bn_act_bwd(dys[i], None, xs[i], mean, invstd, gamma, relu=True, dropout_p=0.3, seed=1, mask=mks[i])),
        "bn_act_bwd_mask_add":  (6.125, lambda i: ops.bn_act_bwd(dys[i], None, xs[i], mean, invstd, gamma, relu=True, mask=mks[i], addend=xs[(i + 1) % nset])),
        "bn_act_bwd":          (7, lambda i: ops.bn_act_bwd(dys[i], ays[i], xs[i], mean, invstd, gamma, relu=True)),
        "bn_act_bwd_dropout":  (7, lambda i: ops.bn_act_bwd(dys[i], ays[i], xs[i], mean, invstd, gamma, relu=True, dropout_p=0.3, seed=1)),
        "bn_act_bwd_addend":   (8, lambda i: ops.bn_act_bwd(dys[i], ays[i], xs[i], mean, invstd, gamma, relu=True, addend=xs[(i + 1) % nset])),
    }
    only = os.environ.get('BENCH_EW_CASES')
    for name, (passes, fn) in cases.items():
        if only and name not in only.split(','):
            continue
        for i in range(nset):
            fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for i in range(nset):
                    fn(i)
        torch.cuda.synchronize()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * nset)
        gbs = passes * tensor_mb * 1e6 / (us * 1e-6) / 1e9
        out.append({"shape": [N, H, H, C], "kernel": name, "us": us, "passes": passes, "GBps": gbs})
        print(f"{H:3d}x{H:<3d}x{C:<4d} {name:22s} {us:7.1f} us  {passes} passes of {tensor_mb:.1f} MB -> {gbs:6.0f} GB/s", flush=True)
json.dump(out, open("gpurun_out/bench_ew%s.json" % os.environ.get("BENCH_TAG", ""), "w"))
