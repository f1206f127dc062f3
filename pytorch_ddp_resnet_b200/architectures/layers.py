"""
Parameter-holding leaf modules and the autograd Functions of the top-level spec tokens.

The modules keep the reference's parameter / buffer names and logical shapes (so state_dicts are
interchangeable with lucaslingle/pytorch_ddp_resnet: resnet/architectures/resnet.py:69-120) but none
of them computes with ATen: their forward passes launch kernels from libb200resnet.so through
pytorch_ddp_resnet_b200.ops. Activations travel between modules as bf16 tensors of logical shape
[N, C, H, W] in channels_last memory format (physically NHWC, the layout the kernels use); conv
filters are fp32 [O, I, kh, kw] parameters in channels_last format (physically KRSC).
"""
import math
from typing import Optional

import torch
import torch as tc

from pytorch_ddp_resnet_b200 import ops
from pytorch_ddp_resnet_b200._lib import B200Error

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------------------
# layout helpers
# --------------------------------------------------------------------------------------------------
def as_nhwc(x: torch.Tensor) -> torch.Tensor:
    """Logical-NCHW activation -> contiguous NHWC bf16 tensor (zero-copy on the fast path)."""
    if x.dim() != 4:
        raise B200Error(f"expected a 4-d activation, got shape {tuple(x.shape)}")
    if not x.is_cuda:
        raise B200Error("pytorch_ddp_resnet_b200 runs on CUDA (sm_100a) only; there is no CPU path")
    v = x.permute(0, 2, 3, 1)
    if ops.get_precision() == "tf32":
        # fp32 / TF32 mode (the reference's un-autocast evaluation): activations stay fp32
        if x.dtype == torch.float32 and v.is_contiguous():
            return v
        if x.dtype == torch.float32 and x.is_contiguous():
            return ops.nchw_f32_to_nhwc_f32(x)
        return v.to(torch.float32).contiguous()
    if x.dtype == torch.bfloat16 and v.is_contiguous():
        return v
    if x.dtype == torch.float32 and x.is_contiguous():
        return ops.nchw_f32_to_nhwc_bf16(x)
    return v.to(torch.bfloat16).contiguous()


def require_forward_only(training: bool) -> None:
    if training:
        raise B200Error("the fp32 / TF32 precision mode is forward-only (evaluation, as the reference uses it: "
                        "evaluation.py:32-39); call .eval() or train in the default bf16 mode")


def conv_forward_with(c, a, w, **kw):
    """Convolution with the geometry of module `c` and an explicit filter `w` ([K,R,S,C], bf16 or fp32 matching
    the activation): used with batch-norm-folded filters in evaluation."""
    if a.dtype == torch.float32:
        return ops.conv_fprop_tf32(a, w, c.stride, c.padding, **kw)
    return ops.conv_fprop(a, w, c.stride, c.padding, **kw)


def conv_forward(c, a, **kw):
    """One convolution of module `c` on activation `a` in the precision the activation carries: bf16 ->
    the bf16 tcgen05 kernels on the cached bf16 filter copy; fp32 -> kind::tf32 on the fp32 master filter."""
    if a.dtype == torch.float32:
        kw.pop("want_stats", None)
        return ops.conv_fprop_tf32(a, c.krsc().contiguous(), c.stride, c.padding, **kw)
    return ops.conv_fprop(a, c.working_copies()[0], c.stride, c.padding, **kw)


def as_nchw_view(x_nhwc: torch.Tensor) -> torch.Tensor:
    """NHWC tensor -> logical NCHW view (channels_last strides)."""
    return x_nhwc.permute(0, 3, 1, 2)


def grad_nhwc(g: torch.Tensor) -> torch.Tensor:
    """Incoming gradient of a logical-NCHW activation -> contiguous NHWC bf16."""
    v = g.permute(0, 2, 3, 1)
    if g.dtype == torch.bfloat16 and v.is_contiguous():
        return v
    return v.to(torch.bfloat16).contiguous()


_dropout_counter = [0]
_weight_generation = [0]


def invalidate_weight_caches() -> None:
    """Marks every cached bf16 filter copy stale (parameters were changed behind torch's back, e.g. by
    a CUDA-graph replay of the optimizer step)."""
    _weight_generation[0] += 1



def next_dropout_seed() -> int:
    """Seeds of the counter-based dropout RNG: torch's global seed mixed with a call counter."""
    _dropout_counter[0] += 1
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + _dropout_counter[0] * 0xD1B54A32D192ED03) % (1 << 64)


def reset_dropout_counter(value: int = 0) -> None:
    _dropout_counter[0] = value


# --------------------------------------------------------------------------------------------------
# parameter holders
# --------------------------------------------------------------------------------------------------
class Conv2d(tc.nn.Module):
    """Holds `weight` [O, I, k, k] (channels_last => KRSC in memory) and an optional `bias` [O]."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, stride: int, padding: int,
                 bias: bool):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding
        w = torch.empty(out_channels, in_channels, kernel_size, kernel_size)
        tc.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        self.weight = tc.nn.Parameter(w.contiguous(memory_format=torch.channels_last))
        if bias:
            bound = 1.0 / math.sqrt(in_channels * kernel_size * kernel_size)
            self.bias = tc.nn.Parameter(torch.empty(out_channels).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)
        self._cache = None
        # optional destinations of the gradients (views into the flat buffer of utils/graph_util.py's
        # bucketed all-reduce): fp32 [K,R,S,C] for the filter, fp32 [K] for the bias
        self.grad_out = None
        self.bias_grad_out = None

    def _apply(self, fn, recurse=True):
        # keep the KRSC physical layout across .to()/.cuda()/.float() (1x1 filters are ambiguous)
        super()._apply(fn, recurse)
        w = self.weight
        if not w.data.permute(0, 2, 3, 1).is_contiguous():
            w.data = w.data.contiguous(memory_format=torch.channels_last)
        self._cache = None
        return self

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, "
                f"stride={self.stride}, padding={self.padding}, bias={self.bias is not None}")

    def krsc(self) -> torch.Tensor:
        """fp32 [K, R, S, C] view of the master weight."""
        v = self.weight.permute(0, 2, 3, 1)
        if not v.is_contiguous():  # e.g. after load_state_dict into a re-strided tensor
            with torch.no_grad():
                self.weight.data = self.weight.data.contiguous(memory_format=torch.channels_last)
            v = self.weight.permute(0, 2, 3, 1)
            if not v.is_contiguous():  # 1x1 filters: strides are ambiguous, memory already is [K,1,1,C]
                v = self.weight.detach().reshape(self.out_channels, self.in_channels)[:, None, None, :]
        return v.detach()

    def _key(self):
        w = self.weight
        return (w.data_ptr(), w._version, _weight_generation[0])

    def working_copies(self):
        """(bf16 KRSC, bf16 CRSK), re-made only when the master weight changed."""
        key = self._key()
        if self._cache is None or self._cache[0] != key:
            wk, wt = ops.weight_prep(self.krsc().contiguous())
            self._cache = (key, wk, wt)
        return self._cache[1], self._cache[2]


class WeightPrepPlan:
    """Batched filter preparation for a whole model: persistent bf16 KRSC / CRSK buffers for every conv
    and ONE kernel launch per step that refreshes all stale ones (instead of two launches per conv).
    prep_subset() refreshes a given group of convs in one launch (the graph-mode step does that per gradient
    bucket, right after the bucket's SGD update, so the next forward finds every copy fresh)."""

    def __init__(self, convs):
        self.convs = list(convs)
        self.table = None
        self.ptrs = None
        self.entries = None
        self._subsets = {}

    def _ensure_buffers(self):
        convs = self.convs
        ptrs = tuple(c.weight.data_ptr() for c in convs)
        if self.table is None or self.ptrs != ptrs:
            entries = []
            for c in convs:
                w = c.krsc().contiguous()
                K, R, S, C = w.shape
                wk = torch.empty((K, R, S, C), dtype=torch.bfloat16, device=w.device)
                wt = torch.empty((C, R, S, K), dtype=torch.bfloat16, device=w.device)
                entries.append((w, wk, wt))
            self.entries = entries
            self.table = ops.weight_prep_table(entries)
            self.ptrs = ptrs
            self._subsets = {}

    def refresh(self):
        convs = self.convs
        if not convs or not convs[0].weight.is_cuda:
            return
        keys = [c._key() for c in convs]
        if all(c._cache is not None and c._cache[0] == k for c, k in zip(convs, keys)):
            return
        self._ensure_buffers()
        ops.weight_prep_multi(self.table, len(convs))
        for c, k, (_, wk, wt) in zip(convs, keys, self.entries):
            c._cache = (k, wk, wt)

    def prep_subset(self, subset):
        """Refreshes the bf16 copies of the convs in `subset` (one launch) and marks them fresh."""
        subset = [c for c in subset if c.weight.is_cuda]
        if not subset:
            return
        self._ensure_buffers()
        index = {id(c): i for i, c in enumerate(self.convs)}
        key = tuple(index[id(c)] for c in subset)
        table = self._subsets.get(key)
        if table is None:
            table = ops.weight_prep_table([self.entries[i] for i in key])
            self._subsets[key] = table
        ops.weight_prep_multi(table, len(key))
        for c, i in zip(subset, key):
            _, wk, wt = self.entries[i]
            c._cache = (c._key(), wk, wt)


def conv_weight_grad(dw_krsc: torch.Tensor) -> torch.Tensor:
    """fp32 [K,R,S,C] kernel output -> gradient shaped like the [O,I,kh,kw] channels_last parameter
    (same strides as the parameter, so autograd / DDP bucket views take it without a copy)."""
    K, R, S, C = dw_krsc.shape
    if R == 1 and S == 1:
        return dw_krsc.view(K, C, 1, 1)
    return dw_krsc.permute(0, 3, 1, 2)


class BatchNorm2d(tc.nn.Module):
    """weight/bias/running_mean/running_var/num_batches_tracked as torch.nn.BatchNorm2d."""

    def __init__(self, num_features: int):
        super().__init__()
        self.num_features = num_features
        self.eps, self.momentum = BN_EPS, BN_MOMENTUM
        self.weight = tc.nn.Parameter(torch.ones(num_features))
        self.bias = tc.nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self.grad_out = None   # optional (dgamma, dbeta) fp32 [C] destinations in a flat gradient buffer

    def grad_dst(self):
        return dict(out_dgamma=self.grad_out[0], out_dbeta=self.grad_out[1]) if self.grad_out else {}

    def batch_stats(self, x_nhwc):
        """Training statistics of x -> (mean, invstd, fwd_kwargs). The running statistics are updated either
        by the statistics launch itself or, when the conv that produced x already computed mean / invstd,
        by the bn_act_fwd launch that must follow with **fwd_kwargs."""
        mean, invstd, deferred = ops.bn_batch_stats(x_nhwc, self.eps, self.momentum, self.running_mean,
                                                    self.running_var, self.num_batches_tracked)
        kw = dict(running=(self.running_mean, self.running_var, self.num_batches_tracked, self.momentum)) \
            if deferred else {}
        return mean, invstd, kw

    def forward(self, x):
        return BnActFn.apply(x, self.weight, self.bias, self, False)

    def extra_repr(self):
        return f"{self.num_features}, eps={self.eps}, momentum={self.momentum}"


class ReLU(tc.nn.Module):
    def forward(self, x):
        return BnActFn.apply(x, None, None, None, True)


class Dropout(tc.nn.Module):
    """Placeholder with the reference's attribute; the blocks fuse dropout into the BN kernels."""

    def __init__(self, p: float):
        super().__init__()
        self.p = p

    def extra_repr(self):
        return f"p={self.p}"


class Linear(tc.nn.Module):
    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = tc.nn.Parameter(torch.empty(out_features, in_features))
        self.bias = tc.nn.Parameter(torch.empty(out_features))
        tc.nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(in_features)
        tc.nn.init.uniform_(self.bias, -bound, bound)
        self.grad_out = None   # optional (dw [O,I], db [O]) destinations in a flat gradient buffer

    def forward(self, x):
        return LinearFn.apply(x, self.weight, self.bias, self)

    def extra_repr(self):
        return f"{self.in_features}, {self.out_features}"


class AvgPool2d(tc.nn.Module):
    def __init__(self, kernel_size: int, stride: int, padding: int):
        super().__init__()
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding

    def forward(self, x):
        return PoolFn.apply(x, self.kernel_size, self.stride, self.padding, False)


class MaxPool2d(tc.nn.Module):
    def __init__(self, kernel_size: int, stride: int, padding: int):
        super().__init__()
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding

    def forward(self, x):
        return PoolFn.apply(x, self.kernel_size, self.stride, self.padding, True)


# --------------------------------------------------------------------------------------------------
# autograd Functions of the top-level tokens
# --------------------------------------------------------------------------------------------------
class TopConvFn(torch.autograd.Function):
    """`cI,O,K,S,P` token: conv with bias (resnet.py:69-75). The 3-channel stem takes the direct path."""

    @staticmethod
    def forward(ctx, x, weight, bias, mod: Conv2d):
        xh = as_nhwc(x)
        if xh.dtype == torch.float32:
            require_forward_only(mod.training)
            return as_nchw_view(conv_forward(mod, xh, bias=bias))
        wk, wt = mod.working_copies()
        # the stem is followed by a batch norm (first block / `n` token): sum its statistics here
        y = ops.conv_fprop(xh, wk, mod.stride, mod.padding, bias=bias, want_stats=mod.training)
        ctx.mod = mod
        ctx.need_dx = x.requires_grad
        ctx.x_dtype, ctx.x_cl = x.dtype, x.permute(0, 2, 3, 1).is_contiguous()
        ctx.save_for_backward(xh, wt)
        return as_nchw_view(y)

    @staticmethod
    def backward(ctx, gy):
        xh, wt = ctx.saved_tensors
        mod = ctx.mod
        g = grad_nhwc(gy)
        k = mod.kernel_size
        dw, db = ops.conv_wgrad(g, xh, k, k, mod.stride, mod.padding, want_dbias=mod.bias is not None,
                                out=mod.grad_out, out_db=mod.bias_grad_out)
        dx = None
        if ctx.need_dx:
            d = ops.conv_dgrad(g, wt, (xh.shape[1], xh.shape[2]), mod.stride, mod.padding)
            dx = as_nchw_view(d)
            if ctx.x_dtype != torch.bfloat16 or not ctx.x_cl:
                dx = dx.to(ctx.x_dtype).contiguous()
        return dx, conv_weight_grad(dw), db, None


class BnActFn(torch.autograd.Function):
    """Top-level `n`, `a` and fused `n a` tokens (resnet.py:111-115)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, bn: Optional[BatchNorm2d], relu: bool):
        xh = as_nhwc(x)
        ctx.relu, ctx.affine, ctx.train = relu, bn is not None, bn is not None and bn.training
        if xh.dtype == torch.float32:
            require_forward_only(bn is not None and bn.training)
        if bn is None:
            y = ops.bn_act_fwd(xh, relu=relu)
            ctx.save_for_backward(y)
        elif bn.training:
            mean, invstd, kw = bn.batch_stats(xh)
            if relu:
                y, mask = ops.bn_act_fwd(xh, mean, invstd, gamma, beta, relu=True, want_mask=True, **kw)
            else:
                y, mask = ops.bn_act_fwd(xh, mean, invstd, gamma, beta, relu=False, **kw), None
            ctx.save_for_backward(xh, mask, mean, invstd, gamma)
            ctx.bn = bn
        else:
            y = ops.bn_act_fwd(xh, bn.running_mean, bn.running_var, gamma, beta, stat_is_var=True,
                               eps=bn.eps, relu=relu)
        return as_nchw_view(y)

    @staticmethod
    def backward(ctx, gy):
        g = grad_nhwc(gy)
        if not ctx.affine:
            (y,) = ctx.saved_tensors
            dx, _, _, _ = ops.bn_act_bwd(g, y, None, relu=ctx.relu)
            return as_nchw_view(dx), None, None, None, None
        if not ctx.train:
            raise B200Error("backward through eval-mode BatchNorm is not supported")
        xh, mask, mean, invstd, gamma = ctx.saved_tensors
        dx, dgamma, dbeta, _ = ops.bn_act_bwd(g, None, xh, mean, invstd, gamma, relu=ctx.relu, mask=mask,
                                              **ctx.bn.grad_dst())
        return as_nchw_view(dx), dgamma, dbeta, None, None


class PoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k: int, stride: int, pad: int, is_max: bool):
        xh = as_nhwc(x)
        ctx.cfg = (k, stride, pad, is_max, tuple(xh.shape))
        if is_max and xh.dtype != torch.float32 and x.requires_grad:
            y, argmax = ops.maxpool_fwd(xh, k, stride, pad, want_argmax=True)
            ctx.save_for_backward(argmax)   # 1 byte per output element; neither x nor y is kept for backward
        else:
            y = ops.maxpool_fwd(xh, k, stride, pad) if is_max else ops.avgpool_fwd(xh, k, stride, pad)
        return as_nchw_view(y)

    @staticmethod
    def backward(ctx, gy):
        k, stride, pad, is_max, shape = ctx.cfg
        g = grad_nhwc(gy)
        if is_max:
            (argmax,) = ctx.saved_tensors
            dx = ops.maxpool_bwd(g, argmax, shape, k, stride, pad)
        else:
            dx = ops.avgpool_bwd(g, shape, k, stride, pad)
        return as_nchw_view(dx), None, None, None, None


class LinearFn(torch.autograd.Function):
    """`fI,O` token after Flatten (resnet.py:117-120): bf16 logits, fp32 parameter gradients."""

    @staticmethod
    def forward(ctx, x, weight, bias, mod=None):
        if not x.is_cuda:
            raise B200Error("pytorch_ddp_resnet_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if x.dtype == torch.float32 and ops.get_precision() == "tf32":
            require_forward_only(mod is not None and mod.training)
            return ops.linear_fwd(x.contiguous(), weight, bias)   # fp32 in, fp32 weights, fp32 logits
        x2 = x if (x.dtype == torch.bfloat16 and x.is_contiguous()) else x.to(torch.bfloat16).contiguous()
        y = ops.linear_fwd(x2, weight, bias)
        ctx.mod = mod
        ctx.need_dx = x.requires_grad
        ctx.x_dtype = x.dtype
        ctx.save_for_backward(x2, weight)
        return y

    @staticmethod
    def backward(ctx, gy):
        x2, weight = ctx.saved_tensors
        g = gy if (gy.dtype == torch.bfloat16 and gy.is_contiguous()) else gy.to(torch.bfloat16).contiguous()
        dst = ctx.mod.grad_out if (ctx.mod is not None and ctx.mod.grad_out) else (None, None)
        dx, dw, db = ops.linear_bwd(g, x2, weight, want_dx=ctx.need_dx, out_dw=dst[0], out_db=dst[1])
        if dx is not None and ctx.x_dtype != torch.bfloat16:
            dx = dx.to(ctx.x_dtype)
        return dx, dw, db, None
