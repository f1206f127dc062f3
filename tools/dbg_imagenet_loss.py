"""Per-step training loss of the ImageNet-shape bench config (bench.py --config wrn50-imagenet) for this repo's
model (eager steps through the C ABI) and for the unmodified reference (oracle/_ref, torch + cuDNN, bf16 autocast),
same initial weights (state_dict copied), same synthetic batches, same SGD. Is the NaN in the bench's last_loss a
divergence of the recipe (lr 0.1, random labels, no warm-up) or a kernel problem?"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

cfg = bench.resolve_config(sys.argv[1] if len(sys.argv) > 1 else "wrn50-imagenet", 1)
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lr = float(sys.argv[3]) if len(sys.argv) > 3 else bench.sgd_args(cfg)["lr"]
B = int(sys.argv[4]) if len(sys.argv) > 4 else cfg["batch"]
hw = cfg["hw"]
dev = torch.device("cuda", 0)
sgd = dict(bench.SGD)
sgd["lr"] = lr

from pytorch_ddp_resnet_b200.algos.metrics import compute_losses_and_metrics  # noqa: E402
from pytorch_ddp_resnet_b200.architectures.resnet import ResNet  # noqa: E402
from pytorch_ddp_resnet_b200.utils.optim_util import get_optimizer  # noqa: E402

torch.manual_seed(0)
ours = ResNet(cfg["spec"], cfg["preact"], cfg["use_proj"], cfg["dropout"]).to(dev).train()
opt = get_optimizer("SGD", ours, dict(sgd))
init = {k: v.clone() for k, v in ours.state_dict().items()}

gen = torch.Generator().manual_seed(1234)
xs = [torch.randn(B, 3, hw, hw, generator=gen).to(dev) for _ in range(4)]
ys = [torch.randint(0, cfg["classes"], (B,), generator=gen).to(dev) for _ in range(4)]

lo = []
for i in range(steps):
    m = compute_losses_and_metrics(logits=ours(xs[i % 4]), labels=ys[i % 4])
    m["loss"].backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    lo.append(m["loss"].item())
del ours, opt
torch.cuda.empty_cache()

lr_ = []
if bench.have_ref():
    sys.path.insert(0, bench.REF_DIR)
    from resnet.architectures.resnet import ResNet as RefResNet
    ref = RefResNet(architecture_spec=cfg["spec"], preact=cfg["preact"], use_proj=cfg["use_proj"],
                    dropout_prob=cfg["dropout"]).to(dev).train()
    ref.load_state_dict(init)
    ropt = torch.optim.SGD(ref.parameters(), **sgd)
    for i in range(steps):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = ref(xs[i % 4])
        loss = torch.nn.functional.cross_entropy(logits.float(), ys[i % 4])
        loss.backward()
        ropt.step()
        ropt.zero_grad(set_to_none=True)
        lr_.append(loss.item())
print("step  ours      reference")
for i in range(steps):
    print(f"{i:3d}  {lo[i]:9.4f}  {lr_[i] if lr_ else float('nan'):9.4f}")
