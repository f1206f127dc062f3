"""
Whole-step CUDA-graph capture: forward + loss/metrics + backward + fused SGD update become ONE graph
launch per step, removing ~200 host-side kernel launches (data parallel: see below).

The training step of the reference (resnet/algos/training.py:94-113) is host-driven op by op; on a
B200 the WRN-28-10 step is a few milliseconds of GPU time, less than the Python time needed to
enqueue it, so the launch-bound inner loop is captured once and replayed:

    step = GraphedTrainStep(classifier, optimizer, x_example, y_example)
    metrics = step(x, y)          # dict of device scalars: loss, top1_err, top5_err

What keeps a replay equal to an eager step:
  * dropout seeds are folded with a device-side step counter that the graph ticks (fresh masks);
  * the learning rate lives in a device scalar that FusedSGD refreshes before each replay, so lr
    schedulers keep working;
  * BN running statistics / num_batches_tracked are updated by kernels inside the graph;
  * cached bf16 filter copies are invalidated after every replay (eval after training sees the
    updated weights).
Inputs must keep the captured shape; other shapes (e.g. a ragged last batch) run eagerly.

Data parallel (classifier is a DistributedDataParallel wrapper): the graph holds forward + backward of
the LOCAL module (no DDP hooks run inside a capture); each replay is followed by ONE NCCL
all-reduce (average) of the flat gradient buffer and the fused SGD launch. BatchNorm buffers stay
per-rank during graphed training and are re-synchronised from rank 0 whenever the DDP wrapper runs a
forward (evaluation, eager steps), which is where the reference's per-forward broadcast is observable.
"""
from typing import Dict, Optional

import torch

from pytorch_ddp_resnet_b200 import ops
from pytorch_ddp_resnet_b200.algos.metrics import compute_losses_and_metrics
from pytorch_ddp_resnet_b200.architectures.layers import invalidate_weight_caches


class GraphedTrainStep:
    def __init__(self, classifier, optimizer, x_example: torch.Tensor, y_example: torch.Tensor,
                 warmup: Optional[int] = None):
        if not x_example.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA tensors")
        self.classifier, self.optimizer = classifier, optimizer
        self.device = x_example.device
        self.static_x = torch.empty_like(x_example).copy_(x_example)
        self.static_y = torch.empty_like(y_example).copy_(y_example)
        self.is_ddp = isinstance(classifier, torch.nn.parallel.DistributedDataParallel)
        self.local = classifier.module if self.is_ddp else classifier
        self.world = torch.distributed.get_world_size() if self.is_ddp else 1
        warmup = 3 if warmup is None else warmup
        self.eager_steps = 0
        self.flat = None
        if self.is_ddp:
            self._setup_flat_gradients()
        ops.step_counter(self.device)  # must exist before capture (an in-capture alloc would re-zero it)
        with torch.cuda.device(self.device):
            ops.bn_accumulators(self.device)  # likewise: zero-filled once, outside the graph
        self.graph = torch.cuda.CUDAGraph()
        classifier.train()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager(self.static_x, self.static_y)
                self.eager_steps += 1
            if hasattr(optimizer, "sync_lr"):
                optimizer.sync_lr()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        from pytorch_ddp_resnet_b200 import _lib
        l0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            ops.tick(self.device)
            self.static_metrics = self._forward_backward(self.local, self.static_x, self.static_y)
            if not self.is_ddp:
                self.optimizer.step()
                self.optimizer.zero_grad(set_to_none=True)
        self.launches_per_step = _lib.launch_count() - l0  # kernels of ours inside one replay
        if self.is_ddp:
            # .grad now aliases the graph's static gradient buffers; every replay rewrites them
            self.grad_owners = [p for p in self.local.parameters() if p.grad is not None]
            self.grads = [p.grad for p in self.grad_owners]
            self.launches_per_step += 1  # the SGD launch that follows each replay
            self._reduce_and_step()      # finish the step that the capture pass computed
        invalidate_weight_caches()

    def _setup_flat_gradients(self) -> None:
        """One flat fp32 buffer holds every parameter gradient, so the exchange step is a single NCCL
        all-reduce. Conv filter gradients (99.9 % of the bytes) are written into it directly by the wgrad
        kernels (Conv2d.grad_out); the small BN / bias / linear gradients are copied in after backward."""
        from pytorch_ddp_resnet_b200.architectures.layers import Conv2d
        params = [p for p in self.local.parameters() if p.requires_grad]
        offsets, total = {}, 0
        for p in params:
            offsets[id(p)] = total
            total += (p.numel() + 3) // 4 * 4  # 16-byte aligned slots
        self.flat = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.flat_views = {}
        for p in params:
            o = offsets[id(p)]
            self.flat_views[id(p)] = self.flat[o:o + p.numel()]
        self.direct = set()
        for m in self.local.modules():
            if isinstance(m, Conv2d):
                K, C, R, S = m.weight.shape
                m.grad_out = self.flat_views[id(m.weight)].view(K, R, S, C)
                self.direct.add(id(m.weight))
        self.flat_params = params

    def _gather_small_grads(self):
        """(flat views, gradient tensors) of the parameters whose kernels do not write into `flat`."""
        dst, src = [], []
        for p in self.flat_params:
            if id(p) in self.direct or p.grad is None:
                continue
            dst.append(self.flat_views[id(p)].view_as(p.grad))
            src.append(p.grad)
        return dst, src

    @staticmethod
    def _forward_backward(module, x, y) -> Dict[str, torch.Tensor]:
        import os
        m = compute_losses_and_metrics(logits=module(x), labels=y)
        # wgrad kernels run on a side stream, concurrently with the dgrad / BN-backward chain
        with ops.wgrad_overlap(x.device, enabled=os.environ.get("B200_WGRAD_OVERLAP", "1") != "0"):
            m["loss"].backward()
        return {k: v.detach() for k, v in m.items()}

    def _exchange(self) -> None:
        """Gradient averaging over ranks: stage the small gradients, then ONE NCCL all-reduce of the flat
        buffer, then point every .grad at its (now averaged) slot."""
        dst, src = self._gather_small_grads()
        if dst:
            torch._foreach_copy_(dst, src)
        torch.distributed.all_reduce(self.flat, op=torch.distributed.ReduceOp.AVG)
        for p in self.flat_params:
            if p.grad is not None and id(p) not in self.direct:
                p.grad = self.flat_views[id(p)].view_as(p.grad)

    def _reduce_and_step(self) -> None:
        # restore the aliases of the graph's static gradient tensors for the staging copy
        for p, g in zip(self.grad_owners, self.grads):
            p.grad = g
        self._exchange()
        self.optimizer.step()

    def _eager(self, x, y) -> Dict[str, torch.Tensor]:
        """One un-captured optimisation step with the same maths as a replay."""
        m = self._forward_backward(self.local, x, y)
        if self.is_ddp:
            self._exchange()
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=True)
        return m

    def matches(self, x: torch.Tensor, y: torch.Tensor) -> bool:
        return (x.shape == self.static_x.shape and x.dtype == self.static_x.dtype
                and y.shape == self.static_y.shape and self.classifier.training)

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> Dict[str, torch.Tensor]:
        """One optimisation step. x / y may live on the host (pinned => asynchronous copy)."""
        if not self.matches(x, y):
            self.classifier.train()
            if self.is_ddp:  # detach the static gradient buffers for the duration of the eager step
                self.optimizer.zero_grad(set_to_none=True)
            out = self._eager(x.to(self.device, non_blocking=True), y.to(self.device, non_blocking=True))
            if self.is_ddp:
                for p, g in zip(self.grad_owners, self.grads):
                    p.grad = g
            return out
        if hasattr(self.optimizer, "sync_lr"):
            self.optimizer.sync_lr()
        self.static_x.copy_(x, non_blocking=True)
        self.static_y.copy_(y, non_blocking=True)
        self.graph.replay()
        if self.is_ddp:
            self._reduce_and_step()
        invalidate_weight_caches()
        return self.static_metrics
