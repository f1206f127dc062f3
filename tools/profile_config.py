"""Eager training steps of one bench.py config, for an ncu launch list:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_config.py NAME [steps]
Prints the number of library launches per step so that the last step can be cut out of the list."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pytorch_ddp_resnet_b200 import _lib  # noqa: E402
from pytorch_ddp_resnet_b200.algos.metrics import compute_losses_and_metrics  # noqa: E402
from pytorch_ddp_resnet_b200.architectures.resnet import ResNet  # noqa: E402
from pytorch_ddp_resnet_b200.utils.optim_util import get_optimizer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "wrn28"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
c = bench.resolve_config(name, 1)
torch.manual_seed(0)
m = ResNet(c["spec"], c["preact"], c["use_proj"], c["dropout"]).cuda().train()
sgd = dict(bench.SGD)
if "lr" in c:
    sgd["lr"] = c["lr"]
opt = get_optimizer("SGD", m, sgd)
x = torch.randn(c["batch"], 3, c["hw"], c["hw"], device="cuda")
y = torch.randint(0, c["classes"], (c["batch"],), device="cuda")
for i in range(steps):
    l0 = _lib.launch_count()
    loss = compute_losses_and_metrics(logits=m(x), labels=y)["loss"]
    loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    print(f"step {i}: {_lib.launch_count() - l0} library launches, loss {loss.item():.4f}", flush=True)
