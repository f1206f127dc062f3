"""Launches each hot kernel a few times at WRN-28-10 shapes (for ncu --set full captures)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_ddp_resnet_b200 import ops, _lib

torch.manual_seed(0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for (N, H, C) in [(128, 32, 160), (128, 16, 320), (128, 8, 640)]:
    x = torch.randn(N, H, H, C, device="cuda").bfloat16()
    dy = torch.randn(N, H, H, C, device="cuda").bfloat16()
    w = (torch.randn(C, 3, 3, C, device="cuda") * 0.02).bfloat16()
    wt = w.permute(3, 1, 2, 0).contiguous()
    gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
    for _ in range(reps):
        y = ops.conv_fprop(x, w, 1, 1, algo=_lib.ALGO_TC)
        y2 = ops.conv_fprop(x, w, 1, 1, residual=dy, algo=_lib.ALGO_TC, want_stats=True)
        mean2, invstd2 = ops.bn_stats(y2, 1e-5)   # finalize of the fused sums
        dx = ops.conv_dgrad(dy, wt, (H, H), 1, 1, algo=_lib.ALGO_TC)
        dw, _ = ops.conv_wgrad(dy, x, 3, 3, 1, 1, algo=_lib.ALGO_TC)
        mean, invstd = ops.bn_stats(x, 1e-5)
        a, mk = ops.bn_act_fwd(x, mean, invstd, gamma, beta, relu=True, dropout_p=0.3, seed=1, want_mask=True)
        g = ops.bn_act_bwd(dy, None, x, mean, invstd, gamma, relu=True, dropout_p=0.3, seed=1, addend=dy, mask=mk)
        # the training path: dgrad with the BN-backward reduction in its epilogue, then the apply pass alone
        da, sums = ops.conv_dgrad_bn_bwd(dy, wt, (H, H), 1, 1, algo=_lib.ALGO_TC, x_bn=x, mask=mk, mean=mean,
                                         invstd=invstd, dropout_p=0.3)
        g2 = ops.bn_act_bwd(da, None, x, mean, invstd, gamma, relu=True, dropout_p=0.3, seed=1, mask=mk, reduced=sums)
torch.cuda.synchronize()
print("done")
