"""
Writes tests/golden/ref_checkpoint/*.pth with the UNMODIFIED reference (imported from /root/reference): a tiny
pre-activation net wrapped in DistributedDataParallel (gloo, world_size 1) is trained for 3 steps by the
reference's own training_loop with a FrequencyCheckpointStrategy, which saves classifier / optimizer /
scheduler / checkpoint_strategy through resnet.utils.checkpoint_util.save_checkpoints. Also stores the batch
and the parameters after ONE MORE reference step (next_step.npz) so that resuming in the new implementation can
be checked against the reference's own continuation. Run in the build container:

    python tests/golden/make_ref_checkpoint.py
"""
import contextlib
import io
import os
import shutil
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, "/root/reference")
from resnet.algos.training import training_loop  # noqa: E402
from resnet.architectures.resnet import ResNet  # noqa: E402
from resnet.utils.checkpoint_util import FrequencyCheckpointStrategy  # noqa: E402
from resnet.utils.optim_util import get_optimizer, get_scheduler  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ref_checkpoint")
SPEC = "c3,16,3,1,1 r1 r1 n a ap16,1,0 fc32,10"
SGD = dict(lr=0.05, momentum=0.9, dampening=0.0, nesterov=True, weight_decay=5e-4)


class _Sampler:
    def set_epoch(self, e):
        pass


def main():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29871")
    dist.init_process_group("gloo", rank=0, world_size=1)
    torch.manual_seed(0)
    model = torch.nn.parallel.DistributedDataParallel(ResNet(SPEC, True, True, 0.0))
    opt = get_optimizer("SGD", model, dict(SGD))
    sched = get_scheduler("MultiStepLR", opt, {"milestones": [2], "gamma": 0.5})
    strat = FrequencyCheckpointStrategy(unit="batch", frequency=3)
    g = torch.Generator().manual_seed(1)
    batches = [(torch.randn(8, 3, 32, 32, generator=g), torch.randint(0, 10, (8,), generator=g)) for _ in range(4)]
    import tempfile
    tmp = tempfile.mkdtemp()   # (the reference's _clean() parses EVERY file name in the checkpoint directory)
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    with contextlib.redirect_stdout(io.StringIO()):
        # steps 0..3: the strategy saves at steps 0 and 3 (files *_1.pth and *_4.pth); we keep step 3's
        training_loop(rank=0, world_size=1, device="cpu", sampler_train=_Sampler(), sampler_test=_Sampler(),
                      dl_train=batches[:3], dl_test=batches[:1], classifier=model, optimizer=opt, scaler=None,
                      scheduler=sched, scheduler_step_unit="batch", checkpoint_strategy=strat,
                      checkpoint_dir=os.path.join(tmp, "ck"), num_microbatches=1, global_step=0, max_steps=3,
                      log_dir=os.path.join(tmp, "tb"))
    from resnet.utils.checkpoint_util import save_checkpoints
    save_checkpoints(OUT, {"checkpoint_strategy": strat, "classifier": model, "optimizer": opt,
                           "scheduler": sched, "scaler": None}, steps=3)
    shutil.rmtree(tmp)
    # the reference's own continuation: one more step on batches[3]
    x, y = batches[3]
    model.train()
    loss = torch.nn.functional.cross_entropy(model(x), y)
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    opt.step()
    np.savez_compressed(os.path.join(OUT, "next_step.npz"), x=x.numpy(), y=y.numpy(), loss=loss.item(),
                        lr=opt.param_groups[0]["lr"],
                        **{"grad/" + n: v.numpy() for n, v in grads.items()},
                        **{"after/" + n: p.detach().numpy() for n, p in model.named_parameters()})
    dist.destroy_process_group()
    print(sorted(os.listdir(OUT)), "lr", opt.param_groups[0]["lr"])


if __name__ == "__main__":
    main()
