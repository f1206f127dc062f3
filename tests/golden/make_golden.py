"""
Generates the golden fixtures in this directory from the UNMODIFIED reference implementation.
Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

For every case: seeded input batch + labels, the reference model's initial state_dict, train-mode
logits / loss / metrics, every parameter gradient, the state after one reference optimizer step
(torch.optim.SGD built by the reference's get_optimizer), and eval-mode logits after that step.
fp32 on CPU; case "wrn_tiny" also stores the logits of the reference under bf16 CPU autocast.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from resnet.architectures.resnet import ResNet  # noqa: E402
from resnet.algos.metrics import compute_losses_and_metrics  # noqa: E402
from resnet.utils.optim_util import get_optimizer  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

SGD = dict(lr=0.1, momentum=0.9, dampening=0.0, nesterov=True, weight_decay=5e-4)

CASES = {
    # reduced ResNet-v1 (non-preact, zero-pad shortcut), BASELINE config 1 family
    "v1_tiny": dict(spec="c3,8,3,1,1 n a r1 r1 r1 ap8,1,0 fc32,10", preact=False, use_proj=False,
                    dropout=0.0, batch=4, hw=32, classes=10),
    # reduced WRN (preact, projection shortcut), BASELINE config 2/5 family
    "wrn_tiny": dict(spec="c3,16,3,1,1 r2 r1 n a ap16,1,0 fc32,10", preact=True, use_proj=True,
                     dropout=0.0, batch=4, hw=32, classes=10),
    # reduced ResNet-v2 bottleneck (preact, projection), BASELINE config 3 family
    "v2_bottleneck_tiny": dict(spec="c3,32,3,1,1 b2 b1 n a ap16,1,0 fc64,10", preact=True,
                               use_proj=True, dropout=0.0, batch=4, hw=32, classes=10),
    # ImageNet-style stem (7x7 s2 + maxpool) with non-preact bottlenecks, BASELINE config 4 family
    "imagenet_style_tiny": dict(spec="c3,32,7,2,3 n a mp3,2,1 b1 b1 ap4,1,0 fc64,12", preact=False,
                                use_proj=True, dropout=0.0, batch=3, hw=32, classes=12),
    # dropout active: masks come from torch's CPU RNG stream (seeded), oracle must consume it alike
    "wrn_dropout_tiny": dict(spec="c3,16,3,1,1 r1 r1 n a ap16,1,0 fc32,10", preact=True, use_proj=True,
                             dropout=0.3, batch=4, hw=32, classes=10),
}


def to_np(d):
    return {k: v.detach().cpu().numpy() for k, v in d.items()}


def main():
    for name, c in CASES.items():
        torch.manual_seed(1234)
        model = ResNet(c["spec"], c["preact"], c["use_proj"], c["dropout"])
        g = torch.Generator().manual_seed(99)
        x = torch.randn(c["batch"], 3, c["hw"], c["hw"], generator=g)
        y = torch.randint(0, c["classes"], (c["batch"],), generator=g)
        init = {k: v.clone() for k, v in model.state_dict().items()}
        out = {"x": x.numpy(), "y": y.numpy()}
        out.update({"init/" + k: v for k, v in to_np(init).items()})

        if name == "wrn_tiny":
            model.train()
            with torch.autocast("cpu", dtype=torch.bfloat16):
                lb = model(x)
            out["bf16_train_logits"] = lb.float().detach().numpy()
            model.load_state_dict(init)

        opt = get_optimizer("SGD", model, dict(SGD))
        model.train()
        torch.manual_seed(4321)  # dropout stream
        logits = model(x)
        metrics = compute_losses_and_metrics(logits=logits, labels=y)
        metrics["loss"].backward()
        out["train_logits"] = logits.detach().numpy()
        for k, v in metrics.items():
            out["metric/" + k] = v.detach().numpy()
        for k, p in model.named_parameters():
            out["grad/" + k] = p.grad.detach().numpy()
        opt.step()
        out.update({"after/" + k: v for k, v in to_np(model.state_dict()).items()})
        model.eval()
        with torch.no_grad():
            out["eval_logits"] = model(x).numpy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "params", sum(p.numel() for p in model.parameters()), "->",
              os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
