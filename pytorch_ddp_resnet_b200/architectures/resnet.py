"""
ResNet(architecture_spec, preact, use_proj, dropout_prob): the public model class of the reference
(resnet/architectures/resnet.py:25-166) rebuilt on the sm_100a kernels.

Same constructor, same spec grammar, same `_architecture.<i>...` module tree and state_dict keys, so
checkpoints move between the two implementations. Differences are all below the API: activations are
bf16 NHWC, every op is a kernel of libb200resnet.so, and adjacent `n a` tokens run as one fused launch.
"""
import re
from typing import List, Tuple

import torch as tc

from pytorch_ddp_resnet_b200 import ops
from pytorch_ddp_resnet_b200.architectures.layers import (
    Conv2d, BatchNorm2d, ReLU, Linear, AvgPool2d, MaxPool2d, TopConvFn, BnActFn, WeightPrepPlan,
    as_nhwc, as_nchw_view, conv_forward_with,
)
from pytorch_ddp_resnet_b200.utils.fold_util import folded_filter
from pytorch_ddp_resnet_b200.architectures.residual_block import (
    ResidualBlock, BottleneckResidualBlock,
)

_TOKEN = re.compile(r"^([a-z]+)((?:\d+)(?:,\d+)*)?$")
_ARITY = {"c": 5, "mp": 3, "ap": 3, "r": 1, "b": 1, "n": 0, "a": 0, "f": 2, "fc": 2}


def tokenize(spec: str) -> List[Tuple[str, Tuple[int, ...]]]:
    """'c3,16,3,1,1 n a r3' -> [('c', (3,16,3,1,1)), ('n', ()), ('a', ()), ('r', (3,))]."""
    out = []
    for tok in spec.split():
        m = _TOKEN.match(tok)
        letters = m.group(1) if m else None
        # the reference dispatches on prefixes, in this order: c, mp, ap, r, b, n, a, f
        kind = next((k for k in ("c", "mp", "ap", "r", "b", "n", "a", "f") if letters and letters.startswith(k)),
                    None)
        if kind is None:
            raise ValueError("Unknown component in architecture spec.")
        args = tuple(int(v) for v in m.group(2).split(",")) if m.group(2) else ()
        if len(args) != _ARITY[kind]:
            raise ValueError(f"component {tok!r} expects {_ARITY[kind]} integers")
        out.append((kind, args))
    return out


class ConvStem(Conv2d):
    """Top-level `c` token: conv WITH bias; forward = TopConvFn."""

    def forward(self, x):
        return TopConvFn.apply(x, self.weight, self.bias, self)


class ResNet(tc.nn.Module):
    def __init__(self, architecture_spec: str, preact: bool, use_proj: bool, dropout_prob: float):
        """
        A residual network.

        :param architecture_spec: space-separated components out of
            {"cI,O,K,S,P", "mpK,S,P", "apK,S,P", "rD", "bD", "n", "a", "fI,O"}:
            convolution (in, out, kernel, stride, padding), max / average pooling (kernel, stride,
            padding), a stack of D basic / bottleneck residual blocks, batch norm, ReLU and
            flatten + fully connected (in, out). A stack that directly follows a stack of the same
            kind downsamples by two and doubles the width in its first block.
        :param preact: Use preactivation ordering?
        :param use_proj: Use projection on skip connection when downsampling?
        :param dropout_prob: Dropout probability.
        """
        super().__init__()
        self._architecture_spec = architecture_spec
        self._preact = preact
        self._use_proj = use_proj
        self._dropout_prob = dropout_prob
        self._architecture = self._build(tokenize(architecture_spec))
        self._init_weights()
        self._prep_plan = None

    def _stack(self, block_cls, depth: int, width_in: int, downsample: bool) -> tc.nn.Sequential:
        width_out = 2 * width_in if downsample else width_in
        blocks = [block_cls(channels=width_in if i == 0 else width_out,
                            downsample=downsample and i == 0, preact=self._preact,
                            use_proj=self._use_proj, dropout_prob=self._dropout_prob)
                  for i in range(depth)]
        return tc.nn.Sequential(*blocks)

    def _build(self, tokens) -> tc.nn.Sequential:
        mods, width, prev = [], None, None
        for kind, a in tokens:
            if kind == "c":
                cin, cout, k, s, p = a
                mods.append(ConvStem(cin, cout, k, s, p, bias=True))
                width = cout
            elif kind == "mp":
                mods.append(MaxPool2d(*a))
            elif kind == "ap":
                mods.append(AvgPool2d(*a))
            elif kind in ("r", "b"):
                down = prev == kind
                cls = ResidualBlock if kind == "r" else BottleneckResidualBlock
                mods.append(self._stack(cls, a[0], width, down))
                width = 2 * width if down else width
            elif kind == "n":
                mods.append(BatchNorm2d(width))
            elif kind == "a":
                mods.append(ReLU())
            elif kind == "f":
                mods.append(tc.nn.Sequential(tc.nn.Flatten(), Linear(*a)))
            prev = kind
        return tc.nn.Sequential(*mods)

    def _init_weights(self):
        # the reference re-initialises only the TOP-LEVEL convs (resnet.py:160-163); block convs keep
        # the default uniform init
        for m in self._architecture:
            if isinstance(m, Conv2d):
                w = tc.empty(m.out_channels, m.in_channels, m.kernel_size, m.kernel_size)
                tc.nn.init.kaiming_normal_(w)
                with tc.no_grad():
                    m.weight.copy_(w)

    def _ensure_prep_plan(self):
        if self._prep_plan is None:
            self._prep_plan = WeightPrepPlan(m for m in self.modules() if isinstance(m, Conv2d))
        return self._prep_plan

    def forward(self, x):
        if x.is_cuda:  # refresh every stale bf16 filter copy of the model in one launch
            self._ensure_prep_plan().refresh()
        mods = list(self._architecture)
        i = 0
        while i < len(mods):
            m = mods[i]
            if (not self.training and ops.get_fold_bn() and isinstance(m, ConvStem) and i + 1 < len(mods)
                    and isinstance(mods[i + 1], BatchNorm2d)):
                # evaluation: a top-level conv directly followed by `n [a]` runs as ONE conv (utils/fold_util.py)
                xh = as_nhwc(x)
                relu = i + 2 < len(mods) and isinstance(mods[i + 2], ReLU)
                w, b = folded_filter(m, mods[i + 1], xh.dtype == tc.float32)
                x = as_nchw_view(conv_forward_with(m, xh, w, bias=b, relu=relu))
                i += 3 if relu else 2
                continue
            if isinstance(m, BatchNorm2d) and i + 1 < len(mods) and isinstance(mods[i + 1], ReLU):
                x = BnActFn.apply(x, m.weight, m.bias, m, True)  # fused `n a`
                i += 2
                continue
            x = m(x)
            i += 1
        return x
