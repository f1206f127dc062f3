"""A few launches of the BN kernels at the WRN shapes (for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_ddp_resnet_b200 import ops
torch.manual_seed(0)
for (N, H, C) in [(128, 32, 160), (128, 8, 640)]:
    x = torch.randn(N, H, H, C, device="cuda").bfloat16()
    dy = torch.randn(N, H, H, C, device="cuda").bfloat16()
    gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
    for _ in range(2):
        mean, invstd = ops.bn_stats(x, 1e-5)
        a = ops.bn_act_fwd(x, mean, invstd, gamma, beta, relu=True, dropout_p=0.3, seed=1)
        g = ops.bn_act_bwd(dy, a, x, mean, invstd, gamma, relu=True, dropout_p=0.3, seed=1, addend=dy)
torch.cuda.synchronize()
print("done")
