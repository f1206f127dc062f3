"""A few launches of the conv kernels at the WRN shapes (for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_ddp_resnet_b200 import ops, _lib
torch.manual_seed(0)
for (N, H, C) in [(128, 32, 160), (128, 16, 320)]:
    x = torch.randn(N, H, H, C, device="cuda").bfloat16()
    dy = torch.randn(N, H, H, C, device="cuda").bfloat16()
    w = (torch.randn(C, 3, 3, C, device="cuda") * 0.02).bfloat16()
    for _ in range(2):
        y = ops.conv_fprop(x, w, 1, 1, algo=_lib.ALGO_TC)
        dw, _ = ops.conv_wgrad(dy, x, 3, 3, 1, 1, algo=_lib.ALGO_TC)
torch.cuda.synchronize()
print("done")
