"""Deterministic mode: where do an eager step, an eager step with wgrad on the side stream and a replay of the
captured step part ways? Prints the tensors that differ after ONE optimisation step from the same state."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_ddp_resnet_b200 import ops  # noqa: E402
from pytorch_ddp_resnet_b200.algos.metrics import compute_losses_and_metrics  # noqa: E402
from pytorch_ddp_resnet_b200.architectures.resnet import ResNet  # noqa: E402
from pytorch_ddp_resnet_b200.utils.graph_util import GraphedTrainStep  # noqa: E402
from pytorch_ddp_resnet_b200.utils.optim_util import get_optimizer  # noqa: E402

SGD = dict(lr=0.1, momentum=0.9, dampening=0.0, nesterov=True, weight_decay=5e-4)
spec = sys.argv[1] if len(sys.argv) > 1 else "c3,16,3,1,1 r1 r1 n a ap16,1,0 fc32,10"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8


def diff(a, b, what):
    bad = [(n, (p.float() - q.float()).abs().max().item()) for (n, p), (_, q) in
           zip(a.state_dict().items(), b.state_dict().items()) if not torch.equal(p, q)]
    print(f"{what}: {len(bad)} of {len(a.state_dict())} tensors differ", bad[:8])


def grads(m):
    return {n: p.grad.clone() for n, p in m.named_parameters()}


def gdiff(a, b, what):
    bad = [(n, (a[n] - b[n]).abs().max().item()) for n in a if not torch.equal(a[n], b[n])]
    print(f"{what}: {len(bad)} of {len(a)} gradients differ", bad[:8])


g = torch.Generator().manual_seed(0)
x = torch.randn(batch, 3, 32, 32, generator=g).cuda()
y = torch.randint(0, 10, (batch,), generator=g).cuda()
torch.manual_seed(0)
base = ResNet(spec, True, True, 0.0).cuda().train()


def fresh():
    m = ResNet(spec, True, True, 0.0).cuda().train()
    m.load_state_dict(base.state_dict())
    return m, get_optimizer("SGD", m, dict(SGD))


with ops.deterministic():
    m1, o1 = fresh()
    compute_losses_and_metrics(logits=m1(x), labels=y)["loss"].backward()
    g1 = grads(m1)
    o1.step()

    m2, o2 = fresh()
    compute_losses_and_metrics(logits=m2(x), labels=y)["loss"].backward()
    gdiff(g1, grads(m2), "eager vs eager gradients")
    o2.step()
    diff(m1, m2, "eager vs eager after the step")

    m3, o3 = fresh()
    loss = compute_losses_and_metrics(logits=m3(x), labels=y)["loss"]
    with ops.wgrad_overlap(x.device, enabled=True):
        loss.backward()
    gdiff(g1, grads(m3), "eager vs eager+wgrad side stream gradients")
    o3.step()
    diff(m1, m3, "eager vs eager+wgrad side stream after the step")

    # second eager step on a model whose momentum buffers exist but are zero (the state a captured step starts from)
    m4, o4 = fresh()
    compute_losses_and_metrics(logits=m4(x), labels=y)["loss"].backward()
    for p in m4.parameters():
        o4.state[p]["momentum_buffer"] = torch.zeros_like(p)
    o4.step()
    diff(m1, m4, "first step (buf = g) vs seasoned step on zero buffers")

    for overlap in ("1", "0"):
        os.environ["B200_WGRAD_OVERLAP"] = overlap
        m5, o5 = fresh()
        step = GraphedTrainStep(m5, o5, x, y)
        diff(base, m5, f"[overlap={overlap}] construction vs initial state")
        step(x, y)
        torch.cuda.synchronize()
        diff(m1, m5, f"[overlap={overlap}] eager vs ONE graph replay")
        m6, o6 = fresh()
        step6 = GraphedTrainStep(m6, o6, x, y)
        step6._eager(x, y)
        torch.cuda.synchronize()
        diff(m1, m6, f"[overlap={overlap}] eager vs GraphedTrainStep._eager")
