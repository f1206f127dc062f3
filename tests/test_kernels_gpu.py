"""
GPU parity of every kernel in libb200resnet.so (called through the C ABI via ctypes) against a plain
torch fp32 reference of the same op on identical bf16-rounded inputs.
Tolerances: bf16 outputs carry one bf16 rounding (2^-9 relative) on top of fp32 accumulation order
=> relative L2 <= 4e-3; fp32 outputs (weight gradients, statistics) <= 1e-3 relative L2 / 1e-4.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.fixture(scope="module", autouse=True)
def _exact_fp32_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _ops():
    from pytorch_ddp_resnet_b200 import ops, _lib
    return ops, _lib


def nhwc(t):  # [N,H,W,C] -> NCHW view for torch
    return t.permute(0, 3, 1, 2)


# N, H, W, C, K, R, stride, pad
CONV_SHAPES = [
    (2, 8, 8, 16, 16, 3, 1, 1),
    (4, 32, 32, 16, 16, 3, 1, 1),
    (4, 16, 16, 32, 64, 3, 1, 1),
    (4, 16, 16, 64, 32, 1, 1, 0),
    (8, 8, 8, 64, 64, 3, 1, 1),
    (4, 32, 32, 160, 160, 3, 1, 1),
    (4, 16, 16, 320, 320, 3, 1, 1),
    (4, 8, 8, 640, 640, 3, 1, 1),
    (4, 32, 32, 160, 320, 3, 2, 1),
    (4, 16, 16, 320, 640, 3, 2, 1),
    (4, 16, 16, 160, 320, 1, 1, 0),
    (2, 32, 32, 16, 32, 3, 2, 1),
    (6, 8, 8, 32, 48, 3, 1, 1),
    (3, 16, 16, 48, 32, 3, 1, 1),
    # ImageNet-style feature maps whose rows do not tile 128 pixels evenly (partial / overhanging tiles)
    (2, 56, 56, 64, 64, 3, 1, 1),
    (3, 28, 28, 64, 128, 1, 1, 0),
    (3, 14, 14, 128, 128, 3, 1, 1),
    (5, 7, 7, 256, 128, 3, 1, 1),
    (2, 28, 28, 64, 64, 3, 2, 1),
]


def _conv_inputs(N, H, W, C, K, R, stride, pad, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(N, H, W, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(K, R, R, C, device="cuda", generator=g) / (C * R * R) ** 0.5).bfloat16()
    P = (H + 2 * pad - R) // stride + 1
    dy = torch.randn(N, P, P, K, device="cuda", generator=g).bfloat16()
    return x, w, dy


@pytest.mark.parametrize("algo", ["tc", "direct"])
@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv_fprop(shape, algo):
    ops, _lib = _ops()
    N, H, W, C, K, R, stride, pad = shape
    a = {"tc": _lib.ALGO_TC, "direct": _lib.ALGO_DIRECT}[algo]
    x, w, _ = _conv_inputs(*shape)
    res = torch.randn(N, (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1, K,
                      device="cuda").bfloat16()
    ref = F.conv2d(nhwc(x).float(), w.permute(0, 3, 1, 2).float(), stride=stride, padding=pad)
    y = ops.conv_fprop(x, w, stride, pad, algo=a)
    assert rel_l2(nhwc(y), ref) < 4e-3
    y2 = ops.conv_fprop(x, w, stride, pad, residual=res, algo=a)
    ref2 = ref.bfloat16().float() + nhwc(res).float()
    assert rel_l2(nhwc(y2), ref2) < 4e-3


@pytest.mark.parametrize("algo", ["tc", "direct"])
@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv_dgrad(shape, algo):
    ops, _lib = _ops()
    N, H, W, C, K, R, stride, pad = shape
    a = {"tc": _lib.ALGO_TC, "direct": _lib.ALGO_DIRECT}[algo]
    x, w, dy = _conv_inputs(*shape)
    w_crsk = w.permute(3, 1, 2, 0).contiguous()
    ref = torch.nn.grad.conv2d_input((N, C, H, W), w.permute(0, 3, 1, 2).float(), nhwc(dy).float(),
                                     stride=stride, padding=pad)
    dx = ops.conv_dgrad(dy, w_crsk, (H, W), stride, pad, algo=a)
    assert rel_l2(nhwc(dx), ref) < 4e-3
    add = torch.randn(N, H, W, C, device="cuda").bfloat16()
    dx2 = ops.conv_dgrad(dy, w_crsk, (H, W), stride, pad, addend=add, algo=a)
    assert rel_l2(nhwc(dx2), ref.bfloat16().float() + nhwc(add).float()) < 4e-3


@pytest.mark.parametrize("algo", ["tc", "direct"])
@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv_wgrad(shape, algo):
    ops, _lib = _ops()
    N, H, W, C, K, R, stride, pad = shape
    a = {"tc": _lib.ALGO_TC, "direct": _lib.ALGO_DIRECT}[algo]
    x, w, dy = _conv_inputs(*shape)
    ref = torch.nn.grad.conv2d_weight(nhwc(x).float(), (K, C, R, R), nhwc(dy).float(), stride=stride,
                                      padding=pad)
    dw, db = ops.conv_wgrad(dy, x, R, R, stride, pad, want_dbias=True, algo=a)
    assert rel_l2(dw.permute(0, 3, 1, 2), ref) < 1e-3
    assert rel_l2(db, dy.float().sum((0, 1, 2))) < 1e-3


@pytest.mark.parametrize("algo", ["auto", "direct"])
@pytest.mark.parametrize("geom", [(8, 32, 160, 3, 1, 1), (4, 32, 32, 7, 2, 3), (2, 224, 64, 7, 2, 3)])
def test_stem_conv_few_input_channels(geom, algo):
    """3-channel stems: 'auto' = im2col + tcgen05 GEMM, 'direct' = CUDA-core kernel."""
    ops, _lib = _ops()
    N, HW, K, R, stride, pad = geom
    a = {"auto": _lib.ALGO_AUTO, "direct": _lib.ALGO_DIRECT}[algo]
    x = torch.randn(N, 3, HW, HW, device="cuda")
    w = torch.randn(K, 3, R, R, device="cuda") * (2.0 / (3 * R * R)) ** 0.5
    b = torch.randn(K, device="cuda") * 0.1
    xh = ops.nchw_f32_to_nhwc_bf16(x)
    assert torch.equal(nhwc(xh), x.bfloat16())
    wk = w.permute(0, 2, 3, 1).contiguous().bfloat16()
    y = ops.conv_fprop(xh, wk, stride, pad, bias=b, algo=a)
    ref = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), b.bfloat16().float(), stride=stride,
                   padding=pad)
    assert rel_l2(nhwc(y), ref) < 4e-3
    dy = torch.randn_like(y)
    dw, db = ops.conv_wgrad(dy, xh, R, R, stride, pad, want_dbias=True, algo=a)
    refw = torch.nn.grad.conv2d_weight(x.bfloat16().float(), (K, 3, R, R), nhwc(dy).float(),
                                       stride=stride, padding=pad)
    assert rel_l2(dw.permute(0, 3, 1, 2), refw) < 1e-3
    assert rel_l2(db, dy.float().sum((0, 1, 2))) < 1e-3


@pytest.mark.parametrize("algo", ["tc", "direct"])
@pytest.mark.parametrize("shape", CONV_SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("with_residual", [False, True])
def test_conv_fprop_fused_bn_statistics(shape, algo, with_residual):
    """conv_fprop(want_stats=True) + bn_stats(y) (finalize only) == conv_fprop + the stand-alone bn_stats,
    for the epilogue-fused kernels (SM pair, halo) and for the accumulate-only fallback (direct)."""
    ops, _lib = _ops()
    N, H, W, C, K, R, stride, pad = shape
    if K % 8:
        pytest.skip("statistics need K % 8 == 0")
    a = {"tc": _lib.ALGO_TC, "direct": _lib.ALGO_DIRECT}[algo]
    if algo == "tc" and not _lib.load().b200_conv2d_tc_supported(_lib.PASS_FPROP, N, H, W, C, K, R, R, stride, pad):
        pytest.skip("shape not on the tcgen05 path")
    x, w, dy = _conv_inputs(*shape)
    res = dy if with_residual else None
    rm, rv = torch.zeros(K, device="cuda"), torch.ones(K, device="cuda")
    nbt = torch.zeros((), dtype=torch.long, device="cuda")
    y = ops.conv_fprop(x, w, stride, pad, residual=res, algo=a, want_stats=True)
    mean, invstd = ops.bn_stats(y, 1e-5, 0.1, rm, rv, nbt)          # consumes the fused sums
    y2 = ops.conv_fprop(x, w, stride, pad, residual=res, algo=a)
    rm2, rv2 = torch.zeros(K, device="cuda"), torch.ones(K, device="cuda")
    mean2, invstd2 = ops.bn_stats(y2, 1e-5, 0.1, rm2, rv2, None)    # stand-alone kernel
    assert torch.equal(y, y2)
    yf = y.float().reshape(-1, K)
    assert (mean - yf.mean(0)).abs().max().item() <= 1e-5 + 1e-4 * yf.abs().mean().item()
    assert rel_l2(mean, mean2) <= 1e-5 or (mean - mean2).abs().max().item() <= 1e-6
    assert rel_l2(invstd, invstd2) <= 1e-5
    assert rel_l2(rm, rm2) <= 1e-5 or (rm - rm2).abs().max().item() <= 1e-6
    assert rel_l2(rv, rv2) <= 1e-5
    assert nbt.item() == 1
    # the accumulators are clean again: an unrelated statistics call is unaffected
    z = torch.randn(64, K, device="cuda").bfloat16()
    m3, _ = ops.bn_stats(z, 1e-5)
    assert (m3 - z.float().mean(0)).abs().max().item() <= 1e-5


def test_unconsumed_fused_statistics_are_discarded():
    """A conv that summed its statistics but is NOT followed by bn_stats on its output must not leak
    those sums into the next user of the accumulators."""
    ops, _lib = _ops()
    x, w, dy = _conv_inputs(4, 16, 16, 64, 64, 3, 1, 1)
    ops.conv_fprop(x, w, 1, 1, want_stats=True)
    z = torch.randn(4, 16, 16, 64, device="cuda").bfloat16()
    m, istd = ops.bn_stats(z, 1e-5)
    zf = z.float().reshape(-1, 64)
    assert (m - zf.mean(0)).abs().max().item() <= 1e-5
    assert rel_l2(istd, (zf.var(0, unbiased=False) + 1e-5).rsqrt()) <= 1e-4
    ops.conv_fprop(x, w, 1, 1, want_stats=True)
    g = ops.bn_act_bwd(dy, None, z, m, istd, torch.ones(64, device="cuda"), relu=False)
    gref = dy.float().reshape(-1, 64).sum(0)
    assert rel_l2(g[2], gref) <= 1e-4


@pytest.mark.parametrize("skip_mode", ["none", "same", "subsample_pad"])
def test_bn_act_mask_path_equals_y_path(skip_mode):
    """bn_act_fwd(want_mask) + bn_act_bwd(mask=...) == the y-based backward, also when the forward added a skip
    tensor before the ReLU (the v1-ordering block tail); deferred running statistics inside bn_act_fwd."""
    ops, _lib = _ops()
    N, H, W, C = 6, 8, 8, 48
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(N, H, W, C, device="cuda", generator=g).bfloat16()
    dy = torch.randn(N, H, W, C, device="cuda", generator=g).bfloat16()
    gamma = torch.rand(C, device="cuda", generator=g) + 0.5
    beta = torch.randn(C, device="cuda", generator=g) * 0.3
    kw = {}
    if skip_mode == "same":
        kw = dict(skip=torch.randn(N, H, W, C, device="cuda", generator=g).bfloat16(), skip_mode=_lib.SKIP_SAME)
    elif skip_mode == "subsample_pad":
        kw = dict(skip=torch.randn(N, 2 * H, 2 * W, C // 2, device="cuda", generator=g).bfloat16(),
                  skip_mode=_lib.SKIP_SUBSAMPLE_PAD)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    nbt = torch.zeros((), dtype=torch.long, device="cuda")
    mean, invstd = ops.bn_stats(x, 1e-5)
    y = ops.bn_act_fwd(x, mean, invstd, gamma, beta, relu=True, **kw)
    ym, mask = ops.bn_act_fwd(x, mean, invstd, gamma, beta, relu=True, want_mask=True,
                              running=(rm, rv, nbt, 0.1), **kw)
    assert torch.equal(y, ym)
    xf = x.float().reshape(-1, C)
    assert nbt.item() == 1
    assert torch.allclose(rm, 0.1 * xf.mean(0), atol=1e-5, rtol=1e-4)
    assert torch.allclose(rv, 0.9 + 0.1 * xf.var(0, unbiased=True), atol=1e-5, rtol=1e-4)
    r1 = ops.bn_act_bwd(dy, y, x, mean, invstd, gamma, relu=True, want_dskip=True)
    r2 = ops.bn_act_bwd(dy, None, x, mean, invstd, gamma, relu=True, want_dskip=True, mask=mask)
    assert torch.equal(r1[3], r2[3])                       # dskip = masked dy
    assert rel_l2(r2[2], r1[2]) < 1e-6 and rel_l2(r2[1], r1[1]) < 1e-5 and rel_l2(r2[0], r1[0]) < 1e-4
    # dropout without ReLU (v1 ordering): the mask is the KEEP mask, also where the input is exactly zero
    xz = torch.relu(x)
    yd, md = ops.bn_act_fwd(xz, relu=False, dropout_p=0.4, seed=7, want_mask=True)
    d1 = ops.bn_act_bwd(dy, None, None, relu=False, dropout_p=0.4, seed=7)[0]
    d2 = ops.bn_act_bwd(dy, None, None, relu=False, dropout_p=0.4, seed=7, mask=md)[0]
    assert torch.equal(d1, d2)
    keep_frac = (d2 != 0).float().mean().item()
    assert abs(keep_frac - 0.6) < 0.03


# N, H (input side), C (input channels = BN channels), K, R, stride, pad, expect the fused epilogue
BNBWD_SHAPES = [
    (8, 32, 160, 160, 3, 1, 1, True),    # halo kernel, 64 + 64 + 32 channel blocks, several tiles per CTA pair
    (16, 16, 320, 320, 3, 1, 1, True),   # halo kernel, two 160-channel output tiles
    (32, 8, 640, 640, 3, 1, 1, True),    # 8x8 images: the SM-pair kernel without halo reuse
    (8, 32, 160, 320, 3, 2, 1, False),   # stride-2 dgrad (four output phases in one launch): left to the reduce pass
    (8, 32, 160, 320, 1, 2, 0, None),    # 1x1 stride 2 (projection): fused or not, the result must agree
    (4, 16, 64, 128, 3, 1, 1, True),
    (2, 8, 24, 40, 3, 1, 1, False),      # channel counts without a tensor-core tile: plain dgrad, sums is None
    (4, 16, 512, 128, 1, 1, 0, True),    # 1x1, reduction length 128 but a small tensor: still fused
    (64, 32, 512, 128, 1, 1, 0, False),  # ... 67 MB of output behind 128 MACs each: the epilogue sums would outlast
                                         # the MMAs, left to the stand-alone reduce pass
    (64, 32, 128, 512, 1, 1, 0, True),   # reduction length 512: fused at any size
]


@pytest.mark.parametrize("p", [0.0, 0.3])
@pytest.mark.parametrize("shape", BNBWD_SHAPES, ids=lambda s: "x".join(map(str, s[:7])))
def test_dgrad_with_fused_bn_backward_reduction(shape, p):
    """conv_dgrad_bn_bwd + bn_act_bwd(reduced=...) == conv_dgrad + bn_act_bwd: the same dx bit for bit out of the conv,
    dgamma / dbeta within fp32 summation order, and the same BN input gradient up to the bf16 rounding that the
    slightly different sums can flip."""
    ops, _lib = _ops()
    N, H, C, K, R, stride, pad, expect = shape
    P = (H + 2 * pad - R) // stride + 1
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(N, H, H, C, device="cuda", generator=g).bfloat16()          # the BN input
    dy = (torch.randn(N, P, P, K, device="cuda", generator=g) * 0.1).bfloat16()  # gradient of the conv output
    w = (torch.randn(K, R, R, C, device="cuda", generator=g) * 0.05)
    _, wt = ops.weight_prep(w)
    gamma = torch.rand(C, device="cuda", generator=g) + 0.5
    beta = torch.randn(C, device="cuda", generator=g) * 0.3
    mean, invstd = ops.bn_stats(x, 1e-5)
    _, mask = ops.bn_act_fwd(x, mean, invstd, gamma, beta, relu=True, dropout_p=p, seed=5, want_mask=True)
    addend = torch.randn(N, H, H, C, device="cuda", generator=g).bfloat16()

    da0 = ops.conv_dgrad(dy, wt, (H, H), stride, pad)
    r0 = ops.bn_act_bwd(da0, None, x, mean, invstd, gamma, relu=True, dropout_p=p, seed=5, mask=mask, addend=addend)
    launches = _lib.launch_count()
    da1, sums = ops.conv_dgrad_bn_bwd(dy, wt, (H, H), stride, pad, x_bn=x, mask=mask, mean=mean, invstd=invstd,
                                      dropout_p=p)
    r1 = ops.bn_act_bwd(da1, None, x, mean, invstd, gamma, relu=True, dropout_p=p, seed=5, mask=mask, addend=addend,
                        reduced=sums)
    if expect is not None:
        assert (sums is not None) == expect
    if sums is not None:
        assert _lib.launch_count() - launches == 2          # the dgrad and the apply pass: no reduce kernel
    assert torch.equal(da0, da1)
    assert rel_l2(r1[2], r0[2]) < 1e-4 and rel_l2(r1[1], r0[1]) < 1e-4   # dbeta, dgamma
    assert rel_l2(r1[0], r0[0]) < 2e-3
    # the accumulators are clean again: the next statistics launch must not see leftovers
    m2, i2 = ops.bn_stats(x, 1e-5)
    assert torch.allclose(m2, mean, atol=1e-6) and torch.allclose(i2, invstd, rtol=1e-5)


def test_weight_prep():
    ops, _ = _ops()
    w = torch.randn(48, 3, 3, 40, device="cuda")
    wk, wt = ops.weight_prep(w)
    assert torch.equal(wk, w.bfloat16())
    assert torch.equal(wt, w.bfloat16().permute(3, 1, 2, 0).contiguous())


@pytest.mark.parametrize("shape", [(4, 32, 32, 160), (8, 8, 8, 640), (3, 5, 7, 16), (2, 4, 4, 4096)])
def test_bn_stats(shape):
    ops, _ = _ops()
    x = (torch.randn(*shape, device="cuda") * 1.7 + 0.4).bfloat16()
    C = shape[-1]
    rm = torch.zeros(C, device="cuda")
    rv = torch.ones(C, device="cuda")
    nbt = torch.zeros((), dtype=torch.int64, device="cuda")
    mean, invstd = ops.bn_stats(x, 1e-5, 0.1, rm, rv, nbt)
    xf = x.float().reshape(-1, C)
    m_ref = xf.mean(0)
    v_ref = xf.var(0, unbiased=False)
    assert torch.allclose(mean, m_ref, atol=1e-4, rtol=1e-4)
    assert torch.allclose(invstd, (v_ref + 1e-5).rsqrt(), atol=1e-4, rtol=1e-4)
    assert torch.allclose(rm, 0.1 * m_ref, atol=1e-4, rtol=1e-4)
    assert torch.allclose(rv, 0.9 + 0.1 * xf.var(0, unbiased=True), atol=1e-4, rtol=1e-4)
    assert nbt.item() == 1


@pytest.mark.parametrize("C", [16, 160, 640])
@pytest.mark.parametrize("relu", [True, False])
def test_bn_act_fwd_bwd_matches_autograd(C, relu):
    ops, _lib = _ops()
    N, H, W = 4, 8, 8
    x = (torch.randn(N, H, W, C, device="cuda") * 1.3 + 0.2).bfloat16()
    gamma = torch.rand(C, device="cuda") + 0.5
    beta = torch.randn(C, device="cuda") * 0.1
    skip = torch.randn(N, H, W, C, device="cuda").bfloat16()
    dy = torch.randn(N, H, W, C, device="cuda").bfloat16()
    mean, invstd = ops.bn_stats(x, 1e-5)
    for use_skip in (False, True):
        y = ops.bn_act_fwd(x, mean, invstd, gamma, beta, relu=relu,
                           skip=skip if use_skip else None, skip_mode=_lib.SKIP_SAME)
        xr = x.float().requires_grad_(True)
        gr = gamma.clone().requires_grad_(True)
        br = beta.clone().requires_grad_(True)
        sr = skip.float().requires_grad_(True)
        t = F.batch_norm(xr.reshape(-1, C), None, None, gr, br, True, 0.1, 1e-5).reshape(N, H, W, C)
        if use_skip:
            t = t + sr
        # teacher-forced gate: the kernel rounds bn(x) and the sum to bf16 like the autocast
        # reference does, so pre-activations within one bf16 ulp of zero may gate differently from
        # this fp32 restatement; use the kernel's own gate for the gradient comparison.
        gate = (y != 0).float()
        ref = t * gate if relu else t
        assert rel_l2(y, torch.relu(t) if relu else t) < 4e-3
        ref.backward(dy.float())
        dx, dgamma, dbeta, dskip = ops.bn_act_bwd(dy, y, x, mean, invstd, gamma, relu=relu,
                                                  want_dskip=use_skip)
        assert rel_l2(dx, xr.grad) < 6e-3
        assert rel_l2(dgamma, gr.grad) < 2e-3
        assert rel_l2(dbeta, br.grad) < 2e-3
        if use_skip:
            assert rel_l2(dskip, sr.grad) < 4e-3


def test_bn_act_eval_mode_and_subsample_pad_skip():
    ops, _lib = _ops()
    N, H, W, C, Cs = 2, 4, 4, 32, 16
    x = torch.randn(N, H, W, C, device="cuda").bfloat16()
    rm = torch.randn(C, device="cuda") * 0.1
    rv = torch.rand(C, device="cuda") + 0.5
    gamma = torch.rand(C, device="cuda") + 0.5
    beta = torch.randn(C, device="cuda") * 0.1
    big = torch.randn(N, 2 * H, 2 * W, Cs, device="cuda").bfloat16()
    y = ops.bn_act_fwd(x, rm, rv, gamma, beta, stat_is_var=True, skip=big,
                       skip_mode=_lib.SKIP_SUBSAMPLE_PAD, relu=True)
    ref = F.batch_norm(x.float().reshape(-1, C), rm, rv, gamma, beta, False, 0.1, 1e-5).reshape(N, H, W, C)
    sk = F.pad(big[:, ::2, ::2, :].float(), (0, C - Cs))
    ref = torch.relu(ref.bfloat16().float() + sk)
    assert rel_l2(y, ref) < 4e-3


def test_dropout_statistics_and_backward_mask():
    ops, _ = _ops()
    x = torch.ones(8, 32, 32, 160, device="cuda").bfloat16()
    p = 0.3
    y = ops.bn_act_fwd(x, relu=False, dropout_p=p, seed=1234)
    keep = (y != 0).float().mean().item()
    assert abs(keep - (1 - p)) < 2e-3
    kept_vals = y[y != 0].float()
    assert torch.allclose(kept_vals, torch.full_like(kept_vals, 1 / (1 - p)), rtol=4e-3)
    y2 = ops.bn_act_fwd(x, relu=False, dropout_p=p, seed=1234)
    assert torch.equal(y, y2)
    y3 = ops.bn_act_fwd(x, relu=False, dropout_p=p, seed=99)
    assert not torch.equal(y, y3)
    dx, _, _, _ = ops.bn_act_bwd(x, None, None, relu=False, dropout_p=p, seed=1234)
    assert torch.equal(dx != 0, y != 0)
    # per-channel keep rate is uniform
    ch = (y != 0).float().mean((0, 1, 2))
    assert (ch - (1 - p)).abs().max().item() < 2e-2


def test_subsample_upsample():
    ops, _ = _ops()
    x = torch.randn(2, 8, 8, 32, device="cuda").bfloat16()
    y = ops.subsample2(x)
    assert torch.equal(y, x[:, ::2, ::2, :])
    dx = torch.randn(2, 8, 8, 32, device="cuda").bfloat16()
    g = torch.randn(2, 4, 4, 16, device="cuda").bfloat16()
    ref = dx.clone().float()
    ref[:, ::2, ::2, :16] += g.float()
    ops.upsample_add_(dx, g)
    assert torch.equal(dx, ref.bfloat16())


@pytest.mark.parametrize("k,s,p,H", [(8, 1, 0, 8), (3, 2, 1, 16), (2, 2, 0, 8),
                                     # compile-time window / stride kernels on odd extents; generic fallback (k = 5)
                                     (3, 2, 1, 15), (3, 2, 0, 17), (3, 1, 1, 12), (2, 2, 0, 14), (5, 2, 2, 16)])
def test_pools(k, s, p, H):
    ops, _ = _ops()
    x = torch.randn(2, H, H, 40 if H % 2 else 32, device="cuda").bfloat16()
    # NCHW-contiguous reference tensors (torch's channels_last avg_pool2d backward differs)
    xa = nhwc(x).float().contiguous().requires_grad_(True)
    ref = F.avg_pool2d(xa, k, s, p)
    y = ops.avgpool_fwd(x, k, s, p)
    assert rel_l2(nhwc(y), ref) < 4e-3
    dy = torch.randn_like(y)
    ref.backward(nhwc(dy).float().contiguous())
    dx = ops.avgpool_bwd(dy, tuple(x.shape), k, s, p)
    assert rel_l2(nhwc(dx), xa.grad) < 8e-3
    xm = nhwc(x).float().contiguous().requires_grad_(True)
    refm = F.max_pool2d(xm, k, s, p)
    ym, am = ops.maxpool_fwd(x, k, s, p, want_argmax=True)
    assert torch.equal(nhwc(ym).float(), refm) and torch.equal(ops.maxpool_fwd(x, k, s, p), ym)
    refm.backward(nhwc(dy).float().contiguous())
    dxm = ops.maxpool_bwd(dy, am, tuple(x.shape), k, s, p)
    assert rel_l2(nhwc(dxm), xm.grad) < 4e-3
    # ties (bf16 inputs collide often): the FIRST maximum in scan order takes the gradient, like torch
    xt = torch.randint(0, 3, x.shape, device="cuda").bfloat16()
    xtm = nhwc(xt).float().contiguous().requires_grad_(True)
    F.max_pool2d(xtm, k, s, p).backward(nhwc(dy).float().contiguous())
    _, amt = ops.maxpool_fwd(xt, k, s, p, want_argmax=True)
    assert rel_l2(nhwc(ops.maxpool_bwd(dy, amt, tuple(xt.shape), k, s, p)), xtm.grad) < 4e-3


def test_maxpool_imagenet_stem_shape_with_ties():
    """k3 s2 p1 over a 112 x 112 map (the ImageNet-style stem pool), values from a 4-symbol alphabet so that nearly
    every window has ties: outputs equal torch's, the gradient lands on the FIRST maximum of every window."""
    ops, _ = _ops()
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randint(0, 4, (4, 112, 112, 64), device="cuda", generator=g).bfloat16()
    xm = nhwc(x).float().contiguous().requires_grad_(True)
    ref = F.max_pool2d(xm, 3, 2, 1)
    y, am = ops.maxpool_fwd(x, 3, 2, 1, want_argmax=True)
    assert torch.equal(nhwc(y).float(), ref)
    dy = torch.randn(4, 56, 56, 64, device="cuda", generator=g).bfloat16()
    ref.backward(nhwc(dy).float().contiguous())
    dx = ops.maxpool_bwd(dy, am, tuple(x.shape), 3, 2, 1)
    assert rel_l2(nhwc(dx), xm.grad) < 4e-3
    assert torch.equal(nhwc(dx).float() != 0, xm.grad != 0)


def test_linear_and_ce():
    ops, _ = _ops()
    B, I, O = 128, 640, 10
    x = torch.randn(B, I, device="cuda").bfloat16()
    w = torch.randn(O, I, device="cuda") * 0.05
    b = torch.randn(O, device="cuda") * 0.1
    y = ops.linear_fwd(x, w, b)
    ref = x.float() @ w.bfloat16().float().t() + b.bfloat16().float()
    assert rel_l2(y, ref) < 4e-3
    labels = torch.randint(0, O, (B,), device="cuda")
    out, dl = ops.ce_topk(y, labels, want_dlogits=True)
    yl = y.float().requires_grad_(True)
    loss = F.cross_entropy(yl, labels)
    loss.backward()
    assert abs(out[0].item() - loss.item()) < 1e-4
    top1 = 1 - (yl.argmax(-1) == labels).float().mean().item()
    top5 = 1 - (yl.topk(5, -1).indices == labels[:, None]).any(-1).float().mean().item()
    assert abs(out[1].item() - top1) < 1e-6 and abs(out[2].item() - top5) < 1e-6
    assert rel_l2(dl, yl.grad) < 4e-3
    dx, dw, db = ops.linear_bwd(dl, x, w)
    assert rel_l2(dx, dl.float() @ w.bfloat16().float()) < 4e-3
    assert rel_l2(dw, dl.float().t() @ x.float()) < 1e-4
    assert rel_l2(db, dl.float().sum(0)) < 1e-4


@pytest.mark.parametrize("nesterov", [False, True])
def test_sgd_matches_torch(nesterov):
    ops, _ = _ops()
    torch.manual_seed(0)
    shapes = [(160, 3, 3, 160), (7,), (640, 10), (1,), (33, 5)]
    ps = [torch.randn(*s, device="cuda") for s in shapes]
    ref_ps = [p.clone().requires_grad_(True) for p in ps]
    opt = torch.optim.SGD(ref_ps, lr=0.1, momentum=0.9, dampening=0.0, nesterov=nesterov,
                          weight_decay=5e-4)
    bufs = [torch.zeros_like(p) for p in ps]
    for step in range(3):
        gs = [torch.randn_like(p) for p in ps]
        for rp, g in zip(ref_ps, gs):
            rp.grad = g.clone()
        opt.step()
        table = torch.tensor([[p.data_ptr() for p in ps], [g.data_ptr() for g in gs],
                              [b.data_ptr() for b in bufs], [p.numel() for p in ps]],
                             dtype=torch.int64).cuda()
        ops.sgd_step(table, len(ps), max(p.numel() for p in ps), 0.1, 0.9, 0.0, 5e-4, nesterov,
                     step == 0)
        for p, rp in zip(ps, ref_ps):
            assert torch.allclose(p, rp.detach(), atol=1e-6, rtol=1e-5)
