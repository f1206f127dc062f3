"""
Loss and metrics (reference: resnet/algos/metrics.py:10-41). Cross entropy (mean), top-1 and top-5
error come out of ONE kernel launch; the backward pass is a second launch of the same kernel that
writes d(loss)/d(logits).
"""
from collections import Counter

import torch as tc

from pytorch_ddp_resnet_b200 import ops
from pytorch_ddp_resnet_b200._lib import B200Error


class _CeTopKFn(tc.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        if not logits.is_cuda:
            raise B200Error("pytorch_ddp_resnet_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if logits.dtype == tc.float32 and ops.get_precision() == "tf32":
            # fp32 / TF32 evaluation: loss and top-k straight from the fp32 logits (forward only)
            out, _ = ops.ce_topk(logits.contiguous(), labels.contiguous(), want_metrics=True)
            ctx.mark_non_differentiable(out)
            return out[0], out[1], out[2]
        lg = logits if (logits.dtype == tc.bfloat16 and logits.is_contiguous()) else logits.to(tc.bfloat16).contiguous()
        out, _ = ops.ce_topk(lg, labels.contiguous(), want_metrics=True)
        ctx.save_for_backward(lg, labels)
        ctx.in_dtype = logits.dtype
        loss, top1, top5 = out[0], out[1], out[2]
        ctx.mark_non_differentiable(top1, top5)
        return loss, top1, top5

    @staticmethod
    def backward(ctx, gloss, _g1, _g5):
        lg, labels = ctx.saved_tensors
        scale = gloss.to(tc.float32).contiguous()
        _, dl = ops.ce_topk(lg, labels.contiguous(), want_metrics=False, want_dlogits=True, grad_scale=scale)
        return dl.to(ctx.in_dtype), None


def cross_entropy_loss(logits, labels):
    return _CeTopKFn.apply(logits, labels)[0]


def top_k_err(logits, labels, k):
    if k not in (1, 5):
        raise ValueError("the fused metric kernel reports top-1 and top-5 error")
    return _CeTopKFn.apply(logits, labels)[1 if k == 1 else 2]


def compute_losses_and_metrics(logits, labels):
    loss, top1_err, top5_err = _CeTopKFn.apply(logits, labels)
    return {
        "loss": loss,
        "top1_err": top1_err,
        "top5_err": top5_err
    }


def global_mean(metric, world_size):
    # for logging purposes only!
    global_metric = metric.clone().float().detach()
    tc.distributed.all_reduce(global_metric, op=tc.distributed.ReduceOp.SUM)
    return global_metric.item() / world_size


def global_means(metrics, world_size):
    """One all-reduce and one host read for all metrics (the reference does one of each per metric,
    metrics.py:32-41); same values."""
    names = list(metrics)
    packed = tc.stack([metrics[k].detach().float().reshape(()) for k in names])
    if tc.distributed.is_available() and tc.distributed.is_initialized():
        tc.distributed.all_reduce(packed, op=tc.distributed.ReduceOp.SUM)
    vals = (packed / world_size).tolist()
    return Counter(dict(zip(names, vals)))


class LaggedMetrics:
    """Metrics without a blocking host read in the step that produced them (reference: metrics.py:32-41 does one
    all-reduce + one `.item()` per metric per microbatch, i.e. three device syncs per step).
    push(): packs the device scalars, all-reduces them once, starts an asynchronous copy into pinned host memory
    and records an event - nothing waits. pop(): returns the values of the OLDEST pushed step once its event has
    completed (by then the device is a whole step further, so the wait is nil)."""

    def __init__(self, world_size: int, depth: int = 2):
        self.world_size, self.depth = world_size, depth
        self.slots = []      # (tag, names, pinned host tensor, event)
        self.free = []

    def push(self, tag, metrics) -> None:
        names = list(metrics)
        packed = tc.stack([metrics[k].detach().float().reshape(()) for k in names])
        if self.world_size > 1 and tc.distributed.is_available() and tc.distributed.is_initialized():
            tc.distributed.all_reduce(packed, op=tc.distributed.ReduceOp.SUM)
        if self.free:
            host, ev = self.free.pop()
        else:
            host = tc.empty(len(names), dtype=tc.float32)
            host = host.pin_memory() if packed.is_cuda else host
            ev = tc.cuda.Event() if packed.is_cuda else None
        host.copy_(packed, non_blocking=True)
        if ev is not None:
            ev.record()
        self.slots.append((tag, names, host, ev))

    def ready(self) -> bool:
        return len(self.slots) >= self.depth

    def pop(self):
        """(tag, Counter of global means) of the oldest pushed step."""
        tag, names, host, ev = self.slots.pop(0)
        if ev is not None:
            ev.synchronize()
        vals = (host / self.world_size).tolist()
        self.free.append((host, ev))
        return tag, Counter(dict(zip(names, vals)))

    def __len__(self):
        return len(self.slots)
