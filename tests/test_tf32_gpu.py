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


# --------------------------------------------------------------------------------------------------
# the evaluation path in the fp32 / TF32 mode (reference: evaluation.py:32-39, no autocast)
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["v1_tiny", "wrn_tiny", "v2_bottleneck_tiny", "imagenet_style_tiny"])
def test_eval_logits_match_fp32_golden_within_1e3(case):
    """Eval-mode logits of the state the UNMODIFIED reference reached after one optimizer step (fp32, CPU)
    against ours in the fp32 / TF32 mode: <= 1e-3 relative L2, identical argmax."""
    from pytorch_ddp_resnet_b200 import ops
    from pytorch_ddp_resnet_b200.architectures.resnet import ResNet
    from tests.golden_util import CASES, load_case
    c, g = CASES[case], load_case(case)
    model = ResNet(c["spec"], c["preact"], c["use_proj"], 0.0)
    model.load_state_dict(g["after"])
    model = model.cuda().eval()
    with torch.no_grad(), ops.precision("tf32"):
        mine = model(g["x"].cuda())
    assert mine.dtype == torch.float32
    ref = g["eval_logits"].cuda()
    e = rel_l2(mine, ref)
    print(f"{case}: tf32-mode eval logits vs fp32 golden rel-L2 {e:.2e}")
    assert e < TF32_TOL
    assert torch.equal(mine.argmax(-1), ref.argmax(-1))
    with pytest.raises(Exception):      # the mode is forward-only
        model.train()
        with ops.precision("tf32"):
            model(g["x"].cuda())


def test_wrn28_10_eval_tf32_mode_batch128():
    """Full WRN-28-10, batch 128, eval mode: kind::tf32 convolutions on every layer (im2col stem, 3x3 s1 / s2, 1x1
    projections) + fp32 BN / pooling / head against the oracle in fp32 on the same GPU, with torch's TF32
    convolutions (the reference's evaluation numerics) and with exact fp32."""
    from oracle import resnet_oracle as O
    from pytorch_ddp_resnet_b200 import ops
    from pytorch_ddp_resnet_b200.algos.metrics import compute_losses_and_metrics
    from pytorch_ddp_resnet_b200.architectures.resnet import ResNet
    spec = "c3,160,3,1,1 r4 r4 r4 n a ap8,1,0 fc640,10"
    init = O.init_state(spec, True, True, seed=3)
    gen = torch.Generator().manual_seed(17)
    for k in init:
        if k.endswith("running_var"):
            init[k] = torch.rand(init[k].shape, generator=gen) + 0.5
        if k.endswith("running_mean"):
            init[k] = torch.randn(init[k].shape, generator=gen) * 0.1
    model = ResNet(spec, True, True, 0.3)
    model.load_state_dict(init)
    model = model.cuda().eval()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(128, 3, 32, 32, generator=g).cuda()
    y = torch.randint(0, 10, (128,), generator=g).cuda()
    state = {k: v.clone().cuda() for k, v in init.items()}
    old = torch.backends.cudnn.allow_tf32
    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = True
        ref_tf32 = O.forward(state, x, spec, True, True, 0.0, training=False)
        torch.backends.cudnn.allow_tf32 = False
        ref_f32 = O.forward(state, x, spec, True, True, 0.0, training=False)
        torch.backends.cudnn.allow_tf32 = old
        with ops.precision("tf32"):
            mine = model(x)
            m = compute_losses_and_metrics(logits=mine, labels=y)
        bf16 = model(x).float()
    e_t, e_f, e_ref, e_b = rel_l2(mine, ref_tf32), rel_l2(mine, ref_f32), rel_l2(ref_tf32, ref_f32), rel_l2(bf16, ref_f32)
    print(f"WRN-28-10 eval logits: ours(tf32 mode) vs torch-TF32 {e_t:.2e}, vs exact fp32 {e_f:.2e}; "
          f"torch-TF32 vs fp32 {e_ref:.2e}; ours(bf16 mode) vs fp32 {e_b:.2e}")
    assert mine.dtype == torch.float32 and e_f < TF32_TOL and e_t < 2 * TF32_TOL
    ref_m = O.losses_and_metrics(ref_f32, y)
    assert abs(m["loss"].item() - ref_m["loss"].item()) < 1e-3
    assert abs(m["top1_err"].item() - ref_m["top1_err"].item()) < 1e-6
    assert torch.equal(mine.argmax(-1), ref_f32.argmax(-1)) or \
        (mine.argmax(-1) != ref_f32.argmax(-1)).float().mean() < 0.02


def test_evaluation_loop_uses_the_reference_precision():
    from pytorch_ddp_resnet_b200.algos.evaluation import evaluation_loop
    from pytorch_ddp_resnet_b200.architectures.resnet import ResNet
    torch.manual_seed(0)
    model = ResNet("c3,32,3,1,1 r1 r1 n a ap16,1,0 fc64,10", True, True, 0.0).cuda()
    g = torch.Generator().manual_seed(2)
    dl = [(torch.randn(16, 3, 32, 32, generator=g), torch.randint(0, 10, (16,), generator=g)) for _ in range(3)]
    a = evaluation_loop(1, "cuda", dl, model)                              # default: tf32, like the reference
    b = evaluation_loop(1, "cuda", dl, model, eval_precision="bf16")
    assert set(a) == {"loss", "top1_err", "top5_err"}
    assert abs(a["loss"] - b["loss"]) < 3e-2 and abs(a["top1_err"] - b["top1_err"]) <= 0.1


# --------------------------------------------------------------------------------------------------
# eval-mode batch-norm folding
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["v1_tiny", "wrn_tiny", "v2_bottleneck_tiny", "imagenet_style_tiny"])
@pytest.mark.parametrize("mode", ["tf32", "bf16"])
def test_bn_folding_preserves_eval_logits(case, mode):
    """`with ops.fold_bn(True)`: conv -> BN(eval) [-> ReLU] pairs run as one conv launch; logits stay within the
    precision mode's tolerance of the unfolded evaluation and of the reference's fp32 golden logits, and fewer
    kernels are launched."""
    from pytorch_ddp_resnet_b200 import _lib, ops
    from pytorch_ddp_resnet_b200.architectures.resnet import ResNet
    from tests.golden_util import CASES, load_case
    c, g = CASES[case], load_case(case)
    model = ResNet(c["spec"], c["preact"], c["use_proj"], 0.0)
    model.load_state_dict(g["after"])
    model = model.cuda().eval()
    x = g["x"].cuda()
    with torch.no_grad(), ops.precision(mode):
        l0 = _lib.launch_count()
        plain = model(x).float()
        l1 = _lib.launch_count()
        with ops.fold_bn(True):
            folded = model(x).float()
        l2 = _lib.launch_count()
    tol = TF32_TOL if mode == "tf32" else 2e-2
    ref = g["eval_logits"].cuda()
    assert rel_l2(folded, plain) < tol and rel_l2(folded, ref) < (tol if mode == "tf32" else 3e-2)
    assert (l2 - l1) < (l1 - l0), "folding must remove launches"
    if mode == "tf32":
        assert torch.equal(folded.argmax(-1), ref.argmax(-1))


def test_bn_folding_wrn28_10_batch128():
    from oracle import resnet_oracle as O
    from pytorch_ddp_resnet_b200 import _lib, ops
    from pytorch_ddp_resnet_b200.architectures.resnet import ResNet
    spec = "c3,160,3,1,1 r4 r4 r4 n a ap8,1,0 fc640,10"
    init = O.init_state(spec, True, True, seed=3)
    gen = torch.Generator().manual_seed(17)
    for k in init:
        if k.endswith("running_var"):
            init[k] = torch.rand(init[k].shape, generator=gen) + 0.5
        if k.endswith("running_mean"):
            init[k] = torch.randn(init[k].shape, generator=gen) * 0.1
    model = ResNet(spec, True, True, 0.3)
    model.load_state_dict(init)
    model = model.cuda().eval()
    x = torch.randn(128, 3, 32, 32, generator=torch.Generator().manual_seed(9)).cuda()
    for mode, tol in (("tf32", TF32_TOL), ("bf16", 2e-2)):
        with torch.no_grad(), ops.precision(mode):
            l0 = _lib.launch_count()
            plain = model(x).float()
            l1 = _lib.launch_count()
            with ops.fold_bn(True):
                folded = model(x).float()
            l2 = _lib.launch_count()
        print(f"WRN-28-10 eval {mode}: folded vs unfolded rel-L2 {rel_l2(folded, plain):.2e}; launches {l1 - l0} -> {l2 - l1}")
        assert rel_l2(folded, plain) < tol
        assert (l1 - l0) - (l2 - l1) >= 12     # conv1 -> norm2 of each of the 12 blocks (+ a one-off filter cast)
