"""
Whole-step CUDA-graph capture: forward + loss/metrics + backward + gradient exchange + fused SGD update
become ONE graph launch per step, removing ~200 host-side kernel launches.

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
Construction is free of side effects: parameters, BN buffers, optimizer state and the dropout step
counter are snapshotted before the warm-up steps that precede a capture and restored after it, so
the first call of the object is optimisation step number one (as in eager mode and in the reference).
Inputs must keep the captured shape; other shapes (e.g. a ragged last batch) run eagerly.

Data parallel (classifier is a DistributedDataParallel wrapper; reference: script.py:64-71, the stock
DDP reducer): the gradient exchange of the reference — bucketed all-reduce (mean) overlapped with the
rest of backward — is rebuilt for graph capture by FlatGradReducer below:
  * ONE flat fp32 buffer holds every parameter gradient, laid out in REVERSE registration order (the
    order in which backward produces them) and cut into ~25 MB buckets at parameter boundaries;
  * every gradient kernel (wgrad, BN backward reduce, linear backward, conv bias) writes straight into
    its slot, so there is no staging copy (a gradient that arrives elsewhere is copied in by the hook);
  * a post-accumulate-grad hook per parameter counts a bucket down; when its last gradient has been
    enqueued, the bucket's ncclAllReduce(avg) is issued on a communication stream that waits for the
    compute stream and the wgrad side stream at that point. Inside a capture those waits become graph
    edges, so each replay overlaps the collectives with the remaining backward kernels;
  * the fused SGD launch waits for the communication stream and is part of the same graph.
BatchNorm buffers stay per-rank during graphed training and are re-synchronised from rank 0 whenever
the DDP wrapper runs a forward (evaluation, eager steps), which is where the reference's per-forward
broadcast is observable.
"""
import os
from typing import Dict, List, Optional

import torch

from pytorch_ddp_resnet_b200 import ops
from pytorch_ddp_resnet_b200.algos.metrics import compute_losses_and_metrics
from pytorch_ddp_resnet_b200.architectures.layers import (BatchNorm2d, Conv2d, Linear,
                                                         invalidate_weight_caches)

DEFAULT_BUCKET_BYTES = 25 << 20


def plan_buckets(sizes_reverse: List[int], bucket_bytes: int = DEFAULT_BUCKET_BYTES,
                 last_bucket_bytes: Optional[int] = None):
    """Cuts a list of gradient sizes (elements, already in reverse registration order) into contiguous
    buckets of about `bucket_bytes`. Returns (offsets, bucket index per entry, [(lo, hi)] per bucket) in
    elements; slots are 16-byte aligned. The bucket that completes last (the first layers of the
    network) is exposed after backward, so it is kept small (`last_bucket_bytes`, default a quarter)."""
    last_bucket_bytes = bucket_bytes // 4 if last_bucket_bytes is None else last_bucket_bytes
    padded = [(n + 3) // 4 * 4 for n in sizes_reverse]
    n = len(padded)
    # tail bucket: the trailing entries (produced last by backward), at least one
    tail_start, acc = n, 0
    while tail_start > 0 and (acc + padded[tail_start - 1]) * 4 <= last_bucket_bytes:
        acc += padded[tail_start - 1]
        tail_start -= 1
    if n and tail_start == n:
        tail_start = n - 1
    offsets, owner, ranges = [], [], []
    lo = pos = 0
    for i, sz in enumerate(padded):
        if i == tail_start and pos > lo:
            ranges.append((lo, pos))
            lo = pos
        offsets.append(pos)
        owner.append(len(ranges))
        pos += sz
        if i < tail_start and (pos - lo) * 4 >= bucket_bytes:
            ranges.append((lo, pos))
            lo = pos
    if pos > lo:
        ranges.append((lo, pos))
    return offsets, owner, ranges


class FlatGradReducer:
    """Flat gradient buffer + bucketed, backward-overlapped NCCL all-reduce (mean) for one module."""

    def __init__(self, module: torch.nn.Module, device: torch.device, world: int,
                 bucket_bytes: int = DEFAULT_BUCKET_BYTES, exchange: bool = True, after_bucket=None):
        """after_bucket(b, params): called on the communication stream right after bucket b's exchange (e.g. the
        optimizer update of exactly those parameters, overlapped with the rest of backward)."""
        self.module, self.device, self.world = module, device, world
        self.exchange = exchange and world > 1
        self.after_bucket = after_bucket
        params = [p for p in module.parameters() if p.requires_grad]
        self.params = params[::-1]
        offsets, owner, self.ranges = plan_buckets([p.numel() for p in self.params], bucket_bytes)
        total = self.ranges[-1][1] if self.ranges else 0
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.slot = {id(p): self.flat[o:o + p.numel()] for p, o in zip(self.params, offsets)}
        self.bucket_of = {id(p): b for p, b in zip(self.params, owner)}
        self.bucket_params = [[] for _ in self.ranges]
        for p, b in zip(self.params, owner):
            self.bucket_params[b].append(p)
        self.bucket_size = [0] * len(self.ranges)
        for b in owner:
            self.bucket_size[b] += 1
        self.pending = list(self.bucket_size)
        self.launched = [False] * len(self.ranges)
        self.active = False
        self.copied = 0   # gradients that had to be copied into their slot (0 on the kernel path)
        self.comm = torch.cuda.Stream(device=device) if device.type == "cuda" else None
        self._attach()
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    def _attach(self) -> None:
        """Points every gradient-producing kernel of the module at its slot of the flat buffer."""
        for m in self.module.modules():
            if isinstance(m, Conv2d):
                K, C, R, S = m.weight.shape
                m.grad_out = self.slot[id(m.weight)].view(K, R, S, C)
                if m.bias is not None:
                    m.bias_grad_out = self.slot[id(m.bias)]
            elif isinstance(m, BatchNorm2d):
                m.grad_out = (self.slot[id(m.weight)], self.slot[id(m.bias)])
            elif isinstance(m, Linear):
                m.grad_out = (self.slot[id(m.weight)].view_as(m.weight), self.slot[id(m.bias)])

    def detach(self) -> None:
        for h in self._hooks:
            h.remove()
        for m in self.module.modules():
            if isinstance(m, Conv2d):
                m.grad_out = m.bias_grad_out = None
            elif isinstance(m, (BatchNorm2d, Linear)):
                m.grad_out = None

    # ---- one backward pass ---------------------------------------------------------------------
    def begin(self) -> None:
        self.pending = list(self.bucket_size)
        self.launched = [False] * len(self.ranges)
        self.active = True

    def _on_grad(self, p: torch.Tensor) -> None:
        if not self.active:
            return
        slot = self.slot[id(p)]
        if p.grad.data_ptr() != slot.data_ptr():
            # a gradient that was not produced in place (a module without grad_out support, or autograd
            # cloned it): move it into the flat buffer so that the exchange and the optimizer see one copy
            dst = slot.as_strided(p.size(), p.stride())
            dst.copy_(p.grad)
            p.grad = dst
            self.copied += 1
        b = self.bucket_of[id(p)]
        self.pending[b] -= 1
        if self.pending[b] == 0:
            self._launch(b)

    def _launch(self, b: int) -> None:
        self.launched[b] = True
        if not self.exchange and self.after_bucket is None:
            return
        lo, hi = self.ranges[b]
        if self.comm is None:   # host tensors (gloo): no streams, no AVG
            if self.exchange:
                torch.distributed.all_reduce(self.flat[lo:hi], op=torch.distributed.ReduceOp.SUM)
                self.flat[lo:hi].div_(self.world)
            if self.after_bucket is not None:
                self.after_bucket(b, self.bucket_params[b])
            return
        cur = torch.cuda.current_stream(self.device)
        self.comm.wait_stream(cur)
        side = ops.wgrad_side_stream_if_any(self.device)
        if side is not None:
            self.comm.wait_stream(side)
        with torch.cuda.stream(self.comm):
            if self.exchange:
                torch.distributed.all_reduce(self.flat[lo:hi], op=torch.distributed.ReduceOp.AVG)
            if self.after_bucket is not None:
                self.after_bucket(b, self.bucket_params[b])

    def finish(self) -> None:
        """After backward: issues the buckets that never filled (parameters without a gradient) and makes
        the current stream wait for the communication stream."""
        self.active = False
        for b in range(len(self.ranges)):
            if not self.launched[b]:
                self._launch(b)
        if (self.exchange or self.after_bucket is not None) and self.comm is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.comm)


class _Snapshot:
    """Parameters, buffers, optimizer state and the dropout step counter of a training setup."""

    def __init__(self, module, optimizer, device):
        self.module, self.optimizer, self.device = module, optimizer, device
        self.tensors = [(t, t.detach().clone()) for t in list(module.parameters()) + list(module.buffers())]
        self.had_state = {id(p): (p in optimizer.state and len(optimizer.state[p]) > 0)
                          for g in optimizer.param_groups for p in g["params"]}
        self.opt_tensors = [(v, v.detach().clone()) for st in optimizer.state.values() for v in st.values()
                            if torch.is_tensor(v)]
        self.counter = ops.step_counter(device).clone()

    def restore(self) -> bool:
        """Returns True when optimizer state was created since the snapshot (it is zeroed, not deleted:
        a captured graph has the 'seasoned' momentum update baked in, and with buf = 0 that update equals
        torch's first step exactly when dampening = 0)."""
        with torch.no_grad():
            for t, c in self.tensors:
                t.copy_(c)
            for t, c in self.opt_tensors:
                t.copy_(c)
            created = False
            for g in self.optimizer.param_groups:
                for p in g["params"]:
                    if not self.had_state.get(id(p), False) and p in self.optimizer.state:
                        for v in self.optimizer.state[p].values():
                            if torch.is_tensor(v):
                                v.zero_()
                                created = True
            ops.step_counter(self.device).copy_(self.counter)
        return created


class GraphedTrainStep:
    def __init__(self, classifier, optimizer, x_example: torch.Tensor, y_example: torch.Tensor,
                 warmup: Optional[int] = None, bucket_bytes: int = DEFAULT_BUCKET_BYTES,
                 exchange: bool = True):
        """exchange=False captures the data-parallel step WITHOUT its collectives (same kernels): the
        difference in step time is the exposed communication (SURVEY 8d timing protocol)."""
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
        # Experiment switch B200_BUCKET_STEP=1: every gradient bucket is exchanged (N > 1), applied by the fused SGD
        # and its filters re-cast to bf16 as soon as backward has produced it, on the communication stream,
        # overlapped with the rest of backward. MEASURED SLOWER on one B200 (5.16 against 5.04 ms/step, round 2): the
        # SGD / cast grids flood every SM with short blocks while a persistent conv kernel (one 200 KB CTA per SM)
        # is trying to get all 148 SMs, so the conv kernels start late. Default: one SGD launch after backward.
        self.bucket_step = hasattr(optimizer, "step_subset") and os.environ.get("B200_BUCKET_STEP", "0") != "0"
        from pytorch_ddp_resnet_b200.architectures.resnet import ResNet
        self.plan = None
        if self.bucket_step and isinstance(self.local, ResNet):
            self.local._ensure_prep_plan()
            self.plan = self.local._prep_plan
        self.reducer = FlatGradReducer(self.local, self.device, self.world, bucket_bytes, exchange,
                                       after_bucket=self._after_bucket if self.bucket_step else None) \
            if (self.is_ddp or self.bucket_step or os.environ.get("B200_FORCE_FLAT", "0") != "0") else None
        self._conv_of = {}
        if self.plan is not None:
            self._conv_of = {id(c.weight): c for c in self.plan.convs}
        ops.step_counter(self.device)  # must exist before capture (an in-capture alloc would re-zero it)
        with torch.cuda.device(self.device):
            ops.bn_accumulators(self.device)  # likewise: zero-filled once, outside the graph
        was_training = classifier.training
        classifier.train()
        snap = _Snapshot(self.local, optimizer, self.device)
        # torch's first momentum step (buf = g) differs from the captured one (buf = m*buf + (1-d)*g on a
        # zeroed buf) only when dampening != 0: such a first step runs eagerly
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):   # >= 1: creates optimizer state and the NCCL communicator
                self._eager(self.static_x, self.static_y)
            if hasattr(optimizer, "sync_lr"):
                optimizer.sync_lr()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        from pytorch_ddp_resnet_b200 import _lib
        l0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            ops.tick(self.device)
            self.static_metrics = self._forward_backward_exchange(self.static_x, self.static_y)
            if not self.bucket_step:
                self.optimizer.step()
            self.optimizer.zero_grad(set_to_none=True)
        self.launches_per_step = _lib.launch_count() - l0  # kernels of ours inside one replay
        torch.cuda.synchronize(self.device)
        created = snap.restore()
        self.eager_first = created and any(g.get("dampening", 0) != 0 and g.get("momentum", 0) != 0
                                           for g in optimizer.param_groups)
        if self.eager_first:
            optimizer.state.clear()
        invalidate_weight_caches()
        self._refresh_filters()   # the restored weights, re-cast into the buffers the graph reads
        if not was_training:
            classifier.eval()

    def _after_bucket(self, b: int, params) -> None:
        """On the communication stream, after bucket b's exchange: SGD update of its parameters, then the bf16
        KRSC / CRSK copies of its conv filters for the NEXT forward (backward is done with this bucket's layers)."""
        self.optimizer.step_subset(params)
        if self.plan is not None:
            convs = [self._conv_of[id(p)] for p in params if id(p) in self._conv_of]
            self.plan.prep_subset(convs)

    def _param_versions(self) -> int:
        return sum(p._version for p in self.local.parameters())

    def _refresh_filters(self) -> None:
        """bf16 filter copies from the current master weights (eagerly, into the persistent buffers). Needed when
        the weights were changed outside the captured step: the graph's forward contains no cast kernel."""
        if self.plan is not None:
            invalidate_weight_caches()
            self.plan.refresh()
        self._versions = self._param_versions()

    def _forward_backward_exchange(self, x, y) -> Dict[str, torch.Tensor]:
        m = compute_losses_and_metrics(logits=self.local(x), labels=y)
        if self.reducer is not None:
            self.reducer.begin()
        # wgrad kernels run on a side stream, concurrently with the dgrad / BN-backward chain
        with ops.wgrad_overlap(x.device, enabled=os.environ.get("B200_WGRAD_OVERLAP", "1") != "0"):
            m["loss"].backward()
        if self.reducer is not None:
            self.reducer.finish()
        return {k: v.detach() for k, v in m.items()}

    def _eager(self, x, y) -> Dict[str, torch.Tensor]:
        """One un-captured optimisation step with the same maths as a replay."""
        m = self._forward_backward_exchange(x, y)
        if not self.bucket_step:
            self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=True)
        return m

    def close(self) -> None:
        """Releases the captured graph (and the NCCL kernels it holds) and detaches the reducer. Call it
        before torch.distributed.destroy_process_group(); the object cannot step afterwards."""
        torch.cuda.synchronize(self.device)
        self.graph.reset()
        self.graph = None
        if self.reducer is not None:
            self.reducer.detach()
            self.reducer = None

    def matches(self, x: torch.Tensor, y: torch.Tensor) -> bool:
        return (x.shape == self.static_x.shape and x.dtype == self.static_x.dtype
                and y.shape == self.static_y.shape and self.classifier.training)

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> Dict[str, torch.Tensor]:
        """One optimisation step. x / y may live on the host (pinned => asynchronous copy)."""
        if not self.matches(x, y) or self.eager_first:
            self.classifier.train()
            self.eager_first = False
            out = self._eager(x.to(self.device, non_blocking=True), y.to(self.device, non_blocking=True))
            invalidate_weight_caches()
            if self.plan is not None:
                self._versions = self._param_versions()   # the eager step re-cast every filter it updated
            return out
        if hasattr(self.optimizer, "sync_lr"):
            self.optimizer.sync_lr()
        if self.plan is not None and self._param_versions() != self._versions:
            self._refresh_filters()   # someone changed the weights between two replays (load_state_dict, ...)
        self.static_x.copy_(x, non_blocking=True)
        self.static_y.copy_(y, non_blocking=True)
        self.graph.replay()
        invalidate_weight_caches()
        return self.static_metrics
