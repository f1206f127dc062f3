"""
Deterministic mode (B200_ALGO_DETERMINISTIC, include/b200resnet.h; `ops.deterministic()` /
torch.use_deterministic_algorithms): SURVEY.md section 8(b) asks the wgrad export for a "deterministic reduction".

Every cross-CTA / cross-warp fp32 reduction then runs in a fixed order, so the SAME inputs give BIT-IDENTICAL
outputs run after run (integer-style bar: torch.equal, no tolerance):
  * wgrad: pixel-range splits store partials, an ordered reduce kernel sums them (halo, general, im2col, direct
    routes; bias gradient) - at the WRN-28-10 batch-128 shapes where the launch really splits;
  * conv epilogues with fused BN statistics / fused BN-backward sums;
  * a whole training step: two runs from the same state agree bit for bit, and replays of the captured CUDA graph
    agree bit for bit with eager steps (small net and WRN-28-10 at batch 128).
The deterministic results stay within fp32 summation-order distance of the default (atomic) ones.
"""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.golden_util import SGD, rel_l2  # noqa: E402

pytestmark = pytest.mark.gpu

REPEATS = 4

# N, H, W, C, K, R, stride, pad
WGRAD_SHAPES = [
    (128, 32, 32, 160, 160, 3, 1, 1),   # halo kernel, ~37 splits
    (128, 16, 16, 320, 320, 3, 1, 1),
    (128, 8, 8, 640, 640, 3, 1, 1),
    (128, 32, 32, 160, 320, 3, 2, 1),   # general kernel, strided windows
    (128, 16, 16, 160, 320, 1, 1, 0),   # 1x1
    (64, 56, 56, 64, 64, 3, 1, 1),      # ImageNet-shape rows
    (4, 16, 16, 32, 64, 3, 1, 1),
]


def _ops():
    from pytorch_ddp_resnet_b200 import ops, _lib
    return ops, _lib


def _inputs(N, H, W, C, K, R, stride, pad, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(N, H, W, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(K, R, R, C, device="cuda", generator=g) / (C * R * R) ** 0.5).bfloat16()
    P = (H + 2 * pad - R) // stride + 1
    dy = torch.randn(N, P, P, K, device="cuda", generator=g).bfloat16()
    return x, w, dy, P


def _perturb():
    """Unrelated work between repeats, so that CTAs of the next launch do not meet the same machine state."""
    a = torch.randn(1 << 22, device="cuda")
    (a * 2).sum()


@pytest.mark.parametrize("shape", WGRAD_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_wgrad_is_bit_reproducible(shape):
    ops, _lib = _ops()
    N, H, W, C, K, R, stride, pad = shape
    x, w, dy, P = _inputs(*shape)
    nws = _lib.load().b200_conv2d_workspace_bytes(_lib.PASS_WGRAD, N, H, W, C, K, R, R, stride, pad,
                                                  _lib.ALGO_AUTO | _lib.ALGO_DETERMINISTIC)
    nws0 = _lib.load().b200_conv2d_workspace_bytes(_lib.PASS_WGRAD, N, H, W, C, K, R, R, stride, pad, _lib.ALGO_AUTO)
    assert nws >= nws0
    if N == 128:
        assert nws - nws0 >= 2 * K * R * R * C * 4, "the batch-128 launches split: partials for >= 2 splits expected"
    dw0, db0 = ops.conv_wgrad(dy, x, R, R, stride, pad, want_dbias=True)        # default: atomics
    outs = []
    with ops.deterministic():
        for _ in range(REPEATS):
            dw, db = ops.conv_wgrad(dy, x, R, R, stride, pad, want_dbias=True)
            outs.append((dw.clone(), db.clone()))
            _perturb()
    for dw, db in outs[1:]:
        assert torch.equal(dw, outs[0][0]), f"dw differs by {(dw - outs[0][0]).abs().max().item():.3e}"
        assert torch.equal(db, outs[0][1]), f"dbias differs by {(db - outs[0][1]).abs().max().item():.3e}"
    assert rel_l2(outs[0][0], dw0) < 1e-5 and rel_l2(outs[0][1], db0) < 1e-5
    # a caller-owned slot (flat gradient bucket) holding garbage: the ordered path must not accumulate into it
    slot = torch.full((K, R, R, C), 7.0, device="cuda")
    with ops.deterministic():
        ops.conv_wgrad(dy, x, R, R, stride, pad, out=slot)
    assert torch.equal(slot, outs[0][0])


@pytest.mark.parametrize("algo", ["direct", "tc"])
def test_wgrad_small_shapes_bit_reproducible_and_correct(algo):
    ops, _lib = _ops()
    a = {"tc": _lib.ALGO_TC, "direct": _lib.ALGO_DIRECT}[algo]
    for shape in [(4, 16, 16, 32, 64, 3, 1, 1), (8, 8, 8, 64, 64, 3, 1, 1), (2, 32, 32, 16, 32, 3, 2, 1),
                  (3, 28, 28, 64, 128, 1, 1, 0)]:
        N, H, W, C, K, R, stride, pad = shape
        x, w, dy, P = _inputs(*shape, seed=3)
        ref = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2).float(), (K, C, R, R), dy.permute(0, 3, 1, 2).float(),
                                          stride=stride, padding=pad)
        with ops.deterministic():
            r = [ops.conv_wgrad(dy, x, R, R, stride, pad, want_dbias=True, algo=a) for _ in range(REPEATS)]
        for dw, db in r[1:]:
            assert torch.equal(dw, r[0][0]) and torch.equal(db, r[0][1])
        assert rel_l2(r[0][0].permute(0, 3, 1, 2), ref) < 1e-3
        assert rel_l2(r[0][1], dy.float().sum((0, 1, 2))) < 1e-3


def test_stem_wgrad_with_bias_is_bit_reproducible():
    """3-channel stem at batch 128: im2col route + the bias gradient over the 42 MB dy (one block per SM)."""
    ops, _lib = _ops()
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(128, 32, 32, 3, device="cuda", generator=g).bfloat16()
    dy = torch.randn(128, 32, 32, 160, device="cuda", generator=g).bfloat16()
    dw0, db0 = ops.conv_wgrad(dy, x, 3, 3, 1, 1, want_dbias=True)
    with ops.deterministic():
        r = []
        for _ in range(REPEATS):
            r.append(ops.conv_wgrad(dy, x, 3, 3, 1, 1, want_dbias=True))
            _perturb()
    for dw, db in r[1:]:
        assert torch.equal(dw, r[0][0]) and torch.equal(db, r[0][1])
    assert rel_l2(r[0][0], dw0) < 1e-5 and rel_l2(r[0][1], db0) < 1e-5
    assert rel_l2(r[0][1], dy.float().sum((0, 1, 2))) < 1e-4


@pytest.mark.parametrize("shape", [(128, 32, 32, 160, 160, 3, 1, 1), (128, 16, 16, 320, 640, 3, 2, 1),
                                   (128, 8, 8, 640, 640, 3, 1, 1), (128, 16, 16, 160, 320, 1, 1, 0)],
                         ids=lambda s: "x".join(map(str, s)))
def test_fused_epilogue_statistics_are_bit_reproducible(shape):
    ops, _lib = _ops()
    N, H, W, C, K, R, stride, pad = shape
    x, w, dy, P = _inputs(*shape, seed=5)
    y0 = ops.conv_fprop(x, w, stride, pad, want_stats=True)
    m0, i0 = ops.bn_stats(y0, 1e-5)
    res = []
    with ops.deterministic():
        for _ in range(REPEATS):
            y = ops.conv_fprop(x, w, stride, pad, want_stats=True)
            m, i = ops.bn_stats(y, 1e-5)
            res.append((y, m.clone(), i.clone()))
            _perturb()
    for y, m, i in res[1:]:
        assert torch.equal(y, res[0][0]) and torch.equal(m, res[0][1]) and torch.equal(i, res[0][2])
    assert torch.equal(res[0][0], y0)
    assert (res[0][1] - m0).abs().max().item() <= 1e-6 and rel_l2(res[0][2], i0) <= 1e-6

    # fused BN-backward sums in the dgrad epilogue (stride 1 only; other shapes return sums = None)
    if stride == 1:
        xin = x                                              # BN input of the layer in front of the conv
        mean, invstd = ops.bn_stats(xin, 1e-5)
        gamma = torch.ones(C, device="cuda")
        beta = torch.zeros(C, device="cuda")
        _, mask = ops.bn_act_fwd(xin, mean, invstd, gamma, beta, relu=True, dropout_p=0.3, seed=9, want_mask=True)
        _, wt = ops.weight_prep(w.float())
        out = []
        with ops.deterministic():
            for _ in range(REPEATS):
                dx, sums = ops.conv_dgrad_bn_bwd(dy, wt, (H, W), stride, pad, x_bn=xin, mask=mask, mean=mean,
                                                 invstd=invstd, dropout_p=0.3)
                if sums is None:     # short reductions (1x1 filters) leave the sums to the stand-alone pass
                    assert R * R * K < 512
                    break
                out.append((dx, sums[0].clone(), sums[1].clone()))
                _perturb()
        for dx, dg, db in out[1:]:
            assert torch.equal(dx, out[0][0]) and torch.equal(dg, out[0][1]) and torch.equal(db, out[0][2])


def test_ce_loss_is_bit_reproducible():
    ops, _ = _ops()
    g = torch.Generator(device="cuda").manual_seed(2)
    for B, O in [(128, 10), (128, 100), (256, 1000), (37, 10)]:
        logits = (torch.randn(B, O, device="cuda", generator=g) * 3).bfloat16()
        labels = torch.randint(0, O, (B,), device="cuda", generator=g)
        outs = [ops.ce_topk(logits, labels, want_dlogits=True) for _ in range(REPEATS)]
        for o, dl in outs[1:]:
            assert torch.equal(o, outs[0][0]) and torch.equal(dl, outs[0][1])
        ref = torch.nn.functional.cross_entropy(logits.float(), labels)
        assert abs(outs[0][0][0].item() - ref.item()) < 1e-3 * max(1.0, abs(ref.item()))
        top1 = (logits.float().argmax(1) != labels).float().mean().item()
        assert abs(outs[0][0][1].item() - top1) < 1e-6


def _train_steps(model, opt, batches):
    from pytorch_ddp_resnet_b200.algos.metrics import compute_losses_and_metrics
    losses = []
    for xb, yb in batches:
        loss = compute_losses_and_metrics(logits=model(xb), labels=yb)["loss"]
        loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
        losses.append(loss.item())
    return losses


def _assert_states_equal(m1, m2, what):
    bad = []
    for (n1, p1), (_, p2) in zip(m1.state_dict().items(), m2.state_dict().items()):
        if not torch.equal(p1, p2):
            bad.append((n1, (p1.float() - p2.float()).abs().max().item()))
    assert not bad, f"{what}: {len(bad)} tensors differ, e.g. {bad[:4]}"


@pytest.mark.parametrize("spec,batch", [("c3,16,3,1,1 r1 r1 n a ap16,1,0 fc32,10", 8),
                                        ("c3,32,3,1,1 r2 r2 r2 n a ap8,1,0 fc128,10", 64)])
def test_training_is_bit_reproducible_and_graph_equals_eager_bitwise(spec, batch):
    from pytorch_ddp_resnet_b200.architectures.resnet import ResNet
    from pytorch_ddp_resnet_b200.utils.graph_util import GraphedTrainStep
    from pytorch_ddp_resnet_b200.utils.optim_util import get_optimizer
    ops, _ = _ops()
    g = torch.Generator().manual_seed(0)
    batches = [(torch.randn(batch, 3, 32, 32, generator=g).cuda(), torch.randint(0, 10, (batch,), generator=g).cuda())
               for _ in range(4)]
    torch.manual_seed(0)
    m1 = ResNet(spec, True, True, 0.0).cuda().train()
    m2 = ResNet(spec, True, True, 0.0).cuda().train()
    m3 = ResNet(spec, True, True, 0.0).cuda().train()
    m2.load_state_dict(m1.state_dict())
    m3.load_state_dict(m1.state_dict())
    o1, o2, o3 = (get_optimizer("SGD", m, dict(SGD)) for m in (m1, m2, m3))
    with ops.deterministic():
        l1 = _train_steps(m1, o1, batches)
        _perturb()
        l2 = _train_steps(m2, o2, batches)
        assert l1 == l2
        _assert_states_equal(m1, m2, "eager run 1 vs eager run 2")
        step = GraphedTrainStep(m3, o3, *batches[0])
        l3 = [step(xb, yb)["loss"].item() for xb, yb in batches]
        assert l3 == l1, (l1, l3)
        _assert_states_equal(m1, m3, "eager vs graph replays")


def test_wrn28_10_batch128_graph_equals_eager_bitwise():
    """At the benchmarked configuration (dropout off: the eager path and the graph draw their masks from different
    step counters): 3 eager steps == 3 replays of the captured step, every parameter and buffer bit for bit."""
    from oracle import resnet_oracle as O
    from pytorch_ddp_resnet_b200.architectures.resnet import ResNet
    from pytorch_ddp_resnet_b200.utils.graph_util import GraphedTrainStep
    from pytorch_ddp_resnet_b200.utils.optim_util import get_optimizer
    ops, _ = _ops()
    spec = "c3,160,3,1,1 r4 r4 r4 n a ap8,1,0 fc640,10"
    init = O.init_state(spec, True, True, seed=1)
    g = torch.Generator().manual_seed(7)
    batches = [(torch.randn(128, 3, 32, 32, generator=g).cuda(), torch.randint(0, 10, (128,), generator=g).cuda())
               for _ in range(3)]
    models = []
    for _ in range(2):
        m = ResNet(spec, True, True, 0.0)
        m.load_state_dict(init)
        models.append(m.cuda().train())
    m1, m2 = models
    o1, o2 = get_optimizer("SGD", m1, dict(SGD)), get_optimizer("SGD", m2, dict(SGD))
    with ops.deterministic():
        l1 = _train_steps(m1, o1, batches)
        step = GraphedTrainStep(m2, o2, *batches[0])
        l2 = [step(xb, yb)["loss"].item() for xb, yb in batches]
    assert l1 == l2, (l1, l2)
    _assert_states_equal(m1, m2, "WRN-28-10 eager vs graph replays")
