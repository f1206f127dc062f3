"""
fp32 / TF32 precision mode (BASELINE.json north_star: "within 1e-3 relative error in fp32/tf32 mode"): the
kind::tf32 convolution kernels against torch's own TF32 convolutions (cuDNN, allow_tf32 = True: what the
reference runs when it evaluates or trains without autocast) and against exact fp32, at the WRN-28-10 shapes
and batch 128. Tolerance 1e-3 relative L2 against the TF32 reference; the distance to exact fp32 is reported
(two correct TF32 implementations sit ~3e-4 .. 8e-4 from fp32).
"""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.golden_util import rel_l2  # noqa: E402

pytestmark = pytest.mark.gpu
TF32_TOL = 1e-3

SHAPES = [  # N, H, W, C, K, R, stride, pad
    (4, 16, 16, 32, 64, 3, 1, 1),
    (8, 8, 8, 64, 64, 3, 1, 1),
    (4, 16, 16, 64, 32, 1, 1, 0),
    (4, 16, 16, 32, 64, 3, 2, 1),
    (128, 32, 32, 160, 160, 3, 1, 1),
    (128, 32, 32, 160, 320, 3, 2, 1),
    (128, 16, 16, 320, 320, 3, 1, 1),
    (128, 16, 16, 160, 320, 1, 1, 0),
    (128, 8, 8, 640, 640, 3, 1, 1),
]


def nhwc(t):
    return t.permute(0, 3, 1, 2)


def _refs(fn):
    """(TF32 result, exact fp32 result) of a torch conv call."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True
    a = fn()
    torch.backends.cudnn.allow_tf32 = False
    b = fn()
    torch.backends.cudnn.allow_tf32 = old
    return a, b


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_tf32_conv_passes(shape):
    from pytorch_ddp_resnet_b200 import _lib, ops
    N, H, W, C, K, R, stride, pad = shape
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, H, W, C, device="cuda", generator=g)
    w = torch.randn(K, R, R, C, device="cuda", generator=g) / (C * R * R) ** 0.5
    P = (H + 2 * pad - R) // stride + 1
    dy = torch.randn(N, P, P, K, device="cuda", generator=g)
    bias = torch.randn(K, device="cuda", generator=g)
    res = torch.randn(N, P, P, K, device="cuda", generator=g)
    wc = w.permute(0, 3, 1, 2)
    report = {}
    # fprop (+ bias + residual, added in fp32 without rounding)
    assert ops.conv_tf32_supported(_lib.PASS_FPROP, N, H, W, C, K, R, R, stride, pad)
    t32, f32 = _refs(lambda: F.conv2d(nhwc(x), wc, stride=stride, padding=pad))
    y = ops.conv_fprop_tf32(x, w, stride, pad)
    report["fprop"] = (rel_l2(nhwc(y), t32), rel_l2(nhwc(y), f32), rel_l2(t32, f32))
    assert report["fprop"][0] < TF32_TOL
    y2 = ops.conv_fprop_tf32(x, w, stride, pad, bias=bias, residual=res)
    assert rel_l2(nhwc(y2), t32 + bias[None, :, None, None] + nhwc(res)) < TF32_TOL
    # dgrad (+ addend)
    if ops.conv_tf32_supported(_lib.PASS_DGRAD, N, H, W, C, K, R, R, stride, pad):
        t32, f32 = _refs(lambda: torch.nn.grad.conv2d_input((N, C, H, W), wc, nhwc(dy), stride=stride, padding=pad))
        w_crsk = w.permute(3, 1, 2, 0).contiguous()
        dx = ops.conv_dgrad_tf32(dy, w_crsk, (H, W), stride, pad)
        report["dgrad"] = (rel_l2(nhwc(dx), t32), rel_l2(nhwc(dx), f32), rel_l2(t32, f32))
        assert report["dgrad"][0] < TF32_TOL
        add = torch.randn(N, H, W, C, device="cuda", generator=g)
        dx2 = ops.conv_dgrad_tf32(dy, w_crsk, (H, W), stride, pad, addend=add)
        assert rel_l2(nhwc(dx2), t32 + nhwc(add)) < TF32_TOL
    # wgrad
    if ops.conv_tf32_supported(_lib.PASS_WGRAD, N, H, W, C, K, R, R, stride, pad):
        t32, f32 = _refs(lambda: torch.nn.grad.conv2d_weight(nhwc(x), (K, C, R, R), nhwc(dy), stride=stride,
                                                             padding=pad))
        dw = ops.conv_wgrad_tf32(dy, x, R, R, stride, pad)
        report["wgrad"] = (rel_l2(dw.permute(0, 3, 1, 2), t32), rel_l2(dw.permute(0, 3, 1, 2), f32), rel_l2(t32, f32))
        assert report["wgrad"][0] < TF32_TOL
    print("tf32", shape, {k: "vs torch-tf32 %.2e, vs fp32 %.2e (torch-tf32 vs fp32 %.2e)" % v for k, v in report.items()})
