"""WRN-28-10 step: GPU ms/step (CUDA events) and host enqueue ms/step (perf_counter, no sync)."""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_ddp_resnet_b200 import _lib
from pytorch_ddp_resnet_b200.algos.metrics import compute_losses_and_metrics
from pytorch_ddp_resnet_b200.architectures.resnet import ResNet
from pytorch_ddp_resnet_b200.utils.optim_util import get_optimizer
import bench
spec = sys.argv[1] if len(sys.argv) > 1 else bench.resolve_config("wrn28", 1)["spec"]
torch.manual_seed(0)
m = ResNet(spec, True, True, 0.3).cuda().train()
opt = get_optimizer("SGD", m, dict(bench.SGD))
x = torch.randn(128, 3, 32, 32, device="cuda"); y = torch.randint(0, 10, (128,), device="cuda")
def step():
    l = compute_losses_and_metrics(logits=m(x), labels=y)["loss"]; l.backward(); opt.step(); opt.zero_grad(set_to_none=True)
for _ in range(5): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
t0 = time.perf_counter(); e0.record()
for _ in range(n): step()
t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
print(json.dumps({"gpu_ms_per_step": e0.elapsed_time(e1) / n, "host_enqueue_ms_per_step": (t1 - t0) * 1e3 / n}))
# forward only / backward only split
torch.cuda.synchronize(); e0.record()
for _ in range(n):
    with torch.no_grad(): m(x)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"fwd_only_ms": e0.elapsed_time(e1) / n}))
