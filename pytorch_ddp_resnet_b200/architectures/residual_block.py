"""
Residual blocks with the constructor signatures, attribute names and state_dict layout of the
reference (resnet/architectures/residual_block.py:8-99 basic, :102-215 bottleneck), executed as ONE
fused autograd node per block:

  pre-activation:  h_j = conv_j( dropout(relu(bn_j(h_{j-1}))) ),  out = h_n + shortcut(x)
                   (the residual add is fused into the last conv's epilogue)
  post-activation: h_j = relu(bn_j(conv_j(dropout(h_{j-1})))),     out = relu(bn_n(conv_n(.)) + shortcut(x))
                   (the add + final ReLU are fused into the last BN kernel)
  shortcut(x):     x | proj_1x1(x[:, :, ::2, ::2]) | zero-channel-pad(x[:, :, ::2, ::2])

The backward pass is written out by hand (dgrad / wgrad / fused BN-ReLU-dropout backward with the
skip gradient folded into the kernels' addend inputs); no ATen op runs on this path.
"""
from typing import List

import torch
import torch as tc

from pytorch_ddp_resnet_b200 import ops, _lib
from pytorch_ddp_resnet_b200._lib import B200Error
from pytorch_ddp_resnet_b200.architectures.layers import (
    Conv2d, BatchNorm2d, ReLU, Dropout, AvgPool2d, as_nhwc, as_nchw_view, grad_nhwc, conv_weight_grad,
    next_dropout_seed, conv_forward, conv_forward_with, require_forward_only,
)
from pytorch_ddp_resnet_b200.utils.fold_util import folded_filter


class _BlockFn(torch.autograd.Function):
    """forward(x, block, *params) with params = [conv weights..., (proj weight), (gamma, beta)...]."""

    @staticmethod
    def forward(ctx, x, block, *params):
        convs: List[Conv2d] = block._convs()
        norms: List[BatchNorm2d] = block._norms()
        n = len(convs)
        proj = block._proj if block._has_proj else None
        gammas = [m.weight for m in norms]
        betas = [m.bias for m in norms]
        p = block._dropout_prob if block.training else 0.0
        training = block.training
        preact = block._preact
        xh = as_nhwc(x)
        f32 = xh.dtype == torch.float32     # fp32 / TF32 precision mode: forward only, fp32 master filters
        if f32:
            require_forward_only(training)
        wk = [(None, None) if f32 else c.working_copies() for c in convs]
        seeds = [next_dropout_seed() if p > 0.0 else 0 for _ in range(n)]

        # ---- shortcut operand -----------------------------------------------------------------
        xsub = None
        if block._downsample and proj is not None:
            xsub = ops.subsample2(xh)
            pt = None if f32 else proj.working_copies()[1]
            skip, skip_mode = conv_forward(proj, xsub), _lib.SKIP_SAME
        elif block._downsample:
            skip, skip_mode = xh, _lib.SKIP_SUBSAMPLE_PAD
        else:
            skip, skip_mode = xh, _lib.SKIP_SAME

        def bn_args(j, t):
            m = norms[j]
            if training:
                mean, invstd, kw = m.batch_stats(t)
                return dict(mean=mean, invstd=invstd, gamma=gammas[j], beta=betas[j], **kw)
            return dict(mean=m.running_mean, invstd=m.running_var, gamma=gammas[j], beta=betas[j],
                        stat_is_var=True, eps=m.eps)

        saved_in, saved_act, saved_conv, stats, masks = [], [], [], [], []
        h = xh
        if not training and ops.get_fold_bn():
            # evaluation with every foldable batch norm merged into the conv in front of it (utils/fold_util.py)
            if preact:
                a = ops.bn_act_fwd(h, relu=True, **bn_args(0, h))
                for j, c in enumerate(convs):
                    if j < n - 1:      # conv_j -> norm_{j+1} -> ReLU in ONE conv launch
                        w, b = folded_filter(c, norms[j + 1], f32)
                        a = conv_forward_with(c, a, w, bias=b, relu=True)
                    elif skip_mode == _lib.SKIP_SAME:
                        out = conv_forward(c, a, residual=skip)
                    else:
                        out = ops.bn_act_fwd(conv_forward(c, a), relu=False, skip=skip, skip_mode=skip_mode)
            else:
                for j, c in enumerate(convs):
                    w, b = folded_filter(c, norms[j], f32)
                    if j < n - 1:
                        h = conv_forward_with(c, h, w, bias=b, relu=True)
                    elif skip_mode == _lib.SKIP_SAME:
                        out = conv_forward_with(c, h, w, bias=b, residual=skip, relu=True)
                    else:
                        out = ops.bn_act_fwd(conv_forward_with(c, h, w, bias=b), relu=True, skip=skip,
                                             skip_mode=skip_mode)
            return as_nchw_view(out)
        if preact:
            for j, c in enumerate(convs):
                st = bn_args(j, h)
                if training:   # the ReLU / dropout bit mask replaces y in the backward kernels
                    a, mk = ops.bn_act_fwd(h, relu=True, dropout_p=p, seed=seeds[j], want_mask=True, **st)
                else:
                    a, mk = ops.bn_act_fwd(h, relu=True, dropout_p=p, seed=seeds[j], **st), None
                masks.append(mk)
                saved_in.append(h)          # BN input
                saved_act.append(a)         # conv input
                stats.append(st)
                # training: the output of these convs feeds a batch norm next (bn j+1 of this block, or
                # the first norm after the block), so its statistics are summed in the conv epilogue
                if j < n - 1:
                    h = conv_forward(c, a, want_stats=training)
                elif skip_mode == _lib.SKIP_SAME:
                    h = conv_forward(c, a, residual=skip, want_stats=training)
                else:
                    h = conv_forward(c, a)
                    h = ops.bn_act_fwd(h, relu=False, skip=skip, skip_mode=skip_mode)
            out = h
        else:
            for j, c in enumerate(convs):
                d = ops.bn_act_fwd(h, relu=False, dropout_p=p, seed=seeds[j]) if p > 0.0 else h
                cj = conv_forward(c, d, want_stats=training)
                st = bn_args(j, cj)
                saved_act.append(d)         # conv input
                saved_conv.append(cj)       # BN input
                stats.append(st)
                kw = dict(want_mask=True) if training else {}
                if j < n - 1:
                    h = ops.bn_act_fwd(cj, relu=True, **st, **kw)
                else:
                    h = ops.bn_act_fwd(cj, relu=True, skip=skip, skip_mode=skip_mode, **st, **kw)
                if training:
                    h, mk = h
                    masks.append(mk)
                saved_in.append(h)          # BN output
            out = h

        ctx.block, ctx.training, ctx.p, ctx.seeds = block, training, p, seeds
        ctx.wt = [w[1] for w in wk]
        ctx.pt = pt if xsub is not None else None
        ctx.xsub = xsub
        ctx.xh = xh
        ctx.saved_in, ctx.saved_act, ctx.saved_conv = saved_in, saved_act, saved_conv
        ctx.masks = masks
        ctx.stats = stats
        ctx.skip_mode = skip_mode
        ctx.need_dx = x.requires_grad
        return as_nchw_view(out)

    @staticmethod
    def backward(ctx, gout):
        block = ctx.block
        if not ctx.training:
            raise B200Error("backward through an eval-mode block is not supported (BN uses running stats)")
        convs: List[Conv2d] = block._convs()
        norms: List[BatchNorm2d] = block._norms()
        n = len(convs)
        p, seeds = ctx.p, ctx.seeds
        identity = not block._downsample
        has_proj = ctx.xsub is not None
        g = grad_nhwc(gout)
        xh = ctx.xh
        dws = [None] * n
        dgs, dbs = [None] * n, [None] * n

        if block._preact:
            dskip = g                      # gradient of the shortcut operand == gradient of the output
            cur = g
            for j in range(n - 1, -1, -1):
                c = convs[j]
                a, hin, st = ctx.saved_act[j], ctx.saved_in[j], ctx.stats[j]
                dw, _ = ops.conv_wgrad(cur, a, c.kernel_size, c.kernel_size, c.stride, c.padding,
                                       out=c.grad_out)
                dws[j] = dw
                # dgrad with the reduction pass of the BN backward in its epilogue (sums is None: no such epilogue)
                da, sums = ops.conv_dgrad_bn_bwd(cur, ctx.wt[j], (a.shape[1], a.shape[2]), c.stride, c.padding,
                                                 x_bn=hin, mask=ctx.masks[j], mean=st["mean"], invstd=st["invstd"],
                                                 dropout_p=p, **norms[j].grad_dst())
                addend = dskip if (j == 0 and identity) else None
                cur, dgs[j], dbs[j], _ = ops.bn_act_bwd(
                    da, None, hin, st["mean"], st["invstd"], st["gamma"], relu=True, dropout_p=p,
                    seed=seeds[j], addend=addend, mask=ctx.masks[j], reduced=sums, **norms[j].grad_dst())
            dx = cur
        else:
            # last BN: relu(bn(c_n) + skip)
            st = ctx.stats[n - 1]
            cur, dgs[n - 1], dbs[n - 1], dskip = ops.bn_act_bwd(
                g, None, ctx.saved_conv[n - 1], st["mean"], st["invstd"], st["gamma"],
                relu=True, want_dskip=True, mask=ctx.masks[n - 1], **norms[n - 1].grad_dst())
            for j in range(n - 1, -1, -1):
                c = convs[j]
                d = ctx.saved_act[j]
                dw, _ = ops.conv_wgrad(cur, d, c.kernel_size, c.kernel_size, c.stride, c.padding,
                                       out=c.grad_out)
                dws[j] = dw
                fuse_skip = (j == 0 and identity and p == 0.0)
                sums = None
                if j > 0 and p == 0.0:
                    # the dgrad output IS the dy of BN j-1's backward: its reduction pass runs in the conv epilogue
                    st = ctx.stats[j - 1]
                    dd, sums = ops.conv_dgrad_bn_bwd(cur, ctx.wt[j], (d.shape[1], d.shape[2]), c.stride, c.padding,
                                                     x_bn=ctx.saved_conv[j - 1], mask=ctx.masks[j - 1],
                                                     mean=st["mean"], invstd=st["invstd"],
                                                     **norms[j - 1].grad_dst())
                else:
                    dd = ops.conv_dgrad(cur, ctx.wt[j], (d.shape[1], d.shape[2]), c.stride, c.padding,
                                        addend=dskip if fuse_skip else None)
                if p > 0.0:  # backward of the dropout in front of conv_j
                    dd, _, _, _ = ops.bn_act_bwd(dd, None, None, relu=False, dropout_p=p, seed=seeds[j],
                                                 addend=dskip if (j == 0 and identity) else None)
                if j > 0:
                    st = ctx.stats[j - 1]
                    cur, dgs[j - 1], dbs[j - 1], _ = ops.bn_act_bwd(
                        dd, None, ctx.saved_conv[j - 1], st["mean"], st["invstd"],
                        st["gamma"], relu=True, mask=ctx.masks[j - 1], reduced=sums, **norms[j - 1].grad_dst())
                else:
                    cur = dd
            dx = cur

        dproj = None
        if has_proj:
            dproj, _ = ops.conv_wgrad(dskip, ctx.xsub, 1, 1, 1, 0, out=block._proj.grad_out)
            dxs = ops.conv_dgrad(dskip, ctx.pt, (ctx.xsub.shape[1], ctx.xsub.shape[2]), 1, 0)
            ops.upsample_add_(dx, dxs)
        elif block._downsample:
            ops.upsample_add_(dx, dskip, Cg=block._in_channels)

        grads = [conv_weight_grad(d) for d in dws]
        if block._has_proj:
            grads.append(conv_weight_grad(dproj))
        for j in range(n):
            grads += [dgs[j], dbs[j]]
        return (as_nchw_view(dx) if ctx.need_dx else None, None, *grads)


class _FusedBlock(tc.nn.Module):
    """Shared machinery of the two block types; subclasses create the reference's submodules."""

    def _convs(self) -> List[Conv2d]:
        raise NotImplementedError

    def _norms(self) -> List[BatchNorm2d]:
        raise NotImplementedError

    @property
    def _has_proj(self) -> bool:
        return self._downsample and self._use_proj

    def forward(self, x):
        params = [c.weight for c in self._convs()]
        if self._has_proj:
            params.append(self._proj.weight)
        for m in self._norms():
            params += [m.weight, m.bias]
        return _BlockFn.apply(x, self, *params)


class ResidualBlock(_FusedBlock):
    def __init__(self, channels: int, downsample: bool, preact: bool, use_proj: bool, dropout_prob: float):
        """
        Basic residual block (two 3x3 convs).

        :param channels: Number of input channels.
        :param downsample: Downsample by a factor of two (and double the width)?
        :param preact: Use preactivation ordering?
        :param use_proj: Use a 1x1 projection on the skip connection when downsampling?
        :param dropout_prob: Dropout probability (applied in front of every conv).
        """
        super().__init__()
        cin = channels
        cout = 2 * channels if downsample else channels
        self._in_channels, self._out_channels = cin, cout
        self._downsample, self._preact, self._use_proj = downsample, preact, use_proj
        self._dropout_prob = dropout_prob
        # registration order == the reference's parameter order (it fixes optimizer / DDP bucket order)
        self._conv1 = Conv2d(cin, cout, 3, 2 if downsample else 1, 1, bias=False)
        self._conv2 = Conv2d(cout, cout, 3, 1, 1, bias=False)
        if downsample:
            self._pool = AvgPool2d(1, 2, 0)
            if use_proj:
                self._proj = Conv2d(cin, cout, 1, 1, 0, bias=False)
        self._norm1 = BatchNorm2d(cin if preact else cout)
        self._norm2 = BatchNorm2d(cout)
        self._act1, self._act2 = ReLU(), ReLU()
        self._dropout1, self._dropout2 = Dropout(dropout_prob), Dropout(dropout_prob)

    def _convs(self):
        return [self._conv1, self._conv2]

    def _norms(self):
        return [self._norm1, self._norm2]


class BottleneckResidualBlock(_FusedBlock):
    def __init__(self, channels: int, downsample: bool, preact: bool, use_proj: bool, dropout_prob: float):
        """
        Bottleneck residual block (1x1 -> 3x3 -> 1x1); the bottleneck is a quarter of the output width.

        :param channels: Number of input channels.
        :param downsample: Downsample by a factor of two (and double the width)?
        :param preact: Use preactivation ordering?
        :param use_proj: Use a 1x1 projection on the skip connection when downsampling?
        :param dropout_prob: Dropout probability (applied in front of every conv).
        """
        super().__init__()
        cin = channels
        cout = 2 * channels if downsample else channels
        mid = channels // 2 if downsample else channels // 4
        self._in_channels, self._bottleneck_channels, self._out_channels = cin, mid, cout
        self._downsample, self._preact, self._use_proj = downsample, preact, use_proj
        self._dropout_prob = dropout_prob
        self._conv1 = Conv2d(cin, mid, 1, 1, 0, bias=False)
        self._conv2 = Conv2d(mid, mid, 3, 2 if downsample else 1, 1, bias=False)
        self._conv3 = Conv2d(mid, cout, 1, 1, 0, bias=False)
        if downsample:
            self._pool = AvgPool2d(1, 2, 0)
            if use_proj:
                self._proj = Conv2d(cin, cout, 1, 1, 0, bias=False)
        self._norm1 = BatchNorm2d(cin if preact else mid)
        self._norm2 = BatchNorm2d(mid)
        self._norm3 = BatchNorm2d(mid if preact else cout)
        self._act1, self._act2, self._act3 = ReLU(), ReLU(), ReLU()
        self._dropout1, self._dropout2, self._dropout3 = (Dropout(dropout_prob), Dropout(dropout_prob),
                                                          Dropout(dropout_prob))

    def _convs(self):
        return [self._conv1, self._conv2, self._conv3]

    def _norms(self):
        return [self._norm1, self._norm2, self._norm3]
