"""
Checkpoint wire-format compatibility (reference: resnet/utils/checkpoint_util.py:52-115): the files
under tests/golden/ref_checkpoint/ were WRITTEN BY THE UNMODIFIED REFERENCE (tests/golden/make_ref_checkpoint.py:
its training_loop + save_checkpoints, DDP-wrapped model, torch.optim.SGD with momentum, MultiStepLR,
FrequencyCheckpointStrategy). They must load into this implementation's objects through its own
maybe_load_checkpoints, and training must continue from them like the reference's own next step.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CK = os.path.join(ROOT, "tests", "golden", "ref_checkpoint")
SPEC = "c3,16,3,1,1 r1 r1 n a ap16,1,0 fc32,10"
SGD = dict(lr=0.05, momentum=0.9, dampening=0.0, nesterov=True, weight_decay=5e-4)


def _build(device):
    from pytorch_ddp_resnet_b200.architectures.resnet import ResNet
    from pytorch_ddp_resnet_b200.utils import checkpoint_util as C
    from pytorch_ddp_resnet_b200.utils.optim_util import get_optimizer, get_scheduler
    model = ResNet(SPEC, True, True, 0.0).to(device)
    if device == "cpu":
        ddp = torch.nn.parallel.DistributedDataParallel(model)
    else:
        from pytorch_ddp_resnet_b200.utils.ddp_util import wrap_ddp
        ddp = wrap_ddp(model, torch.device(device))
    opt = get_optimizer("SGD", ddp, dict(SGD))
    sched = get_scheduler("MultiStepLR", opt, {"milestones": [2], "gamma": 0.5})
    strat = C.get_checkpoint_strategy("FrequencyCheckpointStrategy", {"unit": "batch", "frequency": 3})
    step = C.maybe_load_checkpoints(CK, {"checkpoint_strategy": strat, "classifier": ddp, "optimizer": opt,
                                         "scheduler": sched, "scaler": None}, device, None)
    return model, ddp, opt, sched, strat, step


def _check_loaded(model, opt, sched, strat, step):
    assert step == 3
    sd = torch.load(os.path.join(CK, "classifier_3.pth"), map_location="cpu")
    assert all(k.startswith("module.") for k in sd)
    mine = model.state_dict()
    assert set(mine) == {k[len("module."):] for k in sd}
    for k, v in sd.items():
        assert torch.equal(mine[k[len("module."):]].cpu(), v), k
    assert strat.batch_step == 3 and strat.epoch_step == 1
    assert sched.last_epoch == 3 and abs(opt.param_groups[0]["lr"] - 0.025) < 1e-12
    osd = torch.load(os.path.join(CK, "optimizer_3.pth"), map_location="cpu")
    params = [p for g in opt.param_groups for p in g["params"]]
    assert len(osd["state"]) == len(params)
    for i, p in enumerate(params):   # same parameter order => same optimizer-state indices
        buf = opt.state[p]["momentum_buffer"]
        assert buf.shape == p.shape and torch.equal(buf.cpu(), osd["state"][i]["momentum_buffer"])


def _cpu_worker(rank, port):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=0, world_size=1)
    model, ddp, opt, sched, strat, step = _build("cpu")
    _check_loaded(model, opt, sched, strat, step)
    dist.destroy_process_group()


def test_reference_checkpoint_loads_on_cpu():
    mp.spawn(_cpu_worker, args=(29300 + os.getpid() % 500,), nprocs=1, join=True)


def _gpu_worker(rank, port):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    from pytorch_ddp_resnet_b200.algos.metrics import compute_losses_and_metrics
    model, ddp, opt, sched, strat, step = _build("cuda:0")
    _check_loaded(model, opt, sched, strat, step)
    z = np.load(os.path.join(CK, "next_step.npz"))
    x, y = torch.from_numpy(z["x"]).cuda(), torch.from_numpy(z["y"]).cuda()
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    bufs = {n: opt.state[p]["momentum_buffer"].detach().clone() for n, p in model.named_parameters()}
    ddp.train()
    loss = compute_losses_and_metrics(logits=ddp(x), labels=y)["loss"]
    loss.backward()
    grads = {n: p.grad.detach().clone().float() for n, p in model.named_parameters()}
    opt.step()
    assert abs(loss.item() - float(z["loss"])) < 3e-2      # bf16 forward vs the reference's fp32
    lr, mu, wd = 0.025, SGD["momentum"], SGD["weight_decay"]
    num = den = 0.0
    for n, p in model.named_parameters():
        # exact: the update applied to the LOADED momentum buffer (NCHW-contiguous in the file, re-laid out to the
        # parameter's channels_last strides before the flat-memory kernel touches it) with our own gradient
        g = grads[n].reshape(before[n].shape) + wd * before[n]
        buf = mu * bufs[n] + g
        expect = before[n] - lr * (g + mu * buf)
        assert torch.allclose(p.detach(), expect, atol=1e-6, rtol=1e-5), n
        # and close to the reference's own continuation from the same files
        ref_after = torch.from_numpy(z["after/module." + n]).cuda()
        num += float(((p.detach() - ref_after) ** 2).sum())
        den += float(((ref_after - before[n]) ** 2).sum())
    assert (num / den) ** 0.5 < 0.15, (num / den) ** 0.5   # whole-model update vector, bf16 vs fp32 gradients
    dist.destroy_process_group()


@pytest.mark.gpu
def test_reference_checkpoint_resumes_on_gpu():
    mp.spawn(_gpu_worker, args=(29400 + os.getpid() % 500,), nprocs=1, join=True)
