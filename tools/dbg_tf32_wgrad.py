import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_ddp_resnet_b200 import ops
torch.backends.cudnn.allow_tf32 = False
for (N, H, C, K, R) in [(4, 16, 32, 64, 3), (4, 16, 32, 32, 1), (4, 16, 16, 16, 1)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, H, H, C, device="cuda", generator=g)
    dy = torch.randn(N, H, H, K, device="cuda", generator=g)
    ref = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2), (K, C, R, R), dy.permute(0, 3, 1, 2), padding=R // 2)
    dw = ops.conv_wgrad_tf32(dy, x, R, R, 1, R // 2).permute(0, 3, 1, 2)
    print((N, H, C, K, R), "norms", dw.norm().item(), ref.norm().item(), "nonzero frac", (dw != 0).float().mean().item(),
          "nan", torch.isnan(dw).any().item())
    d = dw.reshape(K, -1); r = ref.reshape(K, -1)
    print(" dw[0,:8]", d[0, :8].tolist()); print(" ref[0,:8]", r[0, :8].tolist())
    # correlation of dw with ref under possible permutations
    print(" cos(dw, ref)", (d * r).sum().item() / (d.norm() * r.norm() + 1e-9).item())
    if R == 1:
        # is dw a permutation of ref entries?
        print(" sorted-abs match", torch.allclose(d.abs().flatten().sort().values, r.abs().flatten().sort().values, rtol=1e-2, atol=1e-2))
