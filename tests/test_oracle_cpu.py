"""
Pins oracle/resnet_oracle.py to the UNMODIFIED reference: every golden fixture (made by
tests/golden/make_golden.py from /root/reference) must be reproduced in fp32 on CPU — logits, loss and
metrics, every parameter gradient, the state after one SGD step (incl. BN running buffers) and the
eval-mode logits. The oracle composes the same torch fp32 ops, so agreement is ~1e-6.
"""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import resnet_oracle as O  # noqa: E402
from tests.golden_util import CASES, SGD, load_case, rel_l2  # noqa: E402


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_reference_fixture(name):
    c, g = CASES[name], load_case(name)
    state = {k: v.clone() for k, v in g["init"].items()}
    # same keys and shapes as the reference state_dict
    ref_init = O.init_state(c["spec"], c["preact"], c["use_proj"])
    assert list(ref_init.keys()) == list(state.keys())
    for k in state:
        assert ref_init[k].shape == state[k].shape and ref_init[k].dtype == state[k].dtype, k
    bufs = {}
    torch.manual_seed(4321)  # the fixture's dropout stream
    out = O.train_step(state, bufs, g["x"], g["y"], c["spec"], c["preact"], c["use_proj"], c["dropout"],
                       dict(SGD))
    assert torch.allclose(out["logits"], g["train_logits"], atol=1e-5, rtol=1e-5)
    for m in ("loss", "top1_err", "top5_err"):
        assert abs(out[m].item() - g["metric"][m].item()) < 1e-6, m
    assert set(out["grads"]) == set(g["grad"])
    for k, v in g["grad"].items():
        assert torch.allclose(out["grads"][k], v, atol=1e-6, rtol=1e-4), k
    for k, v in g["after"].items():
        if v.dtype == torch.int64:
            assert torch.equal(state[k], v), k
        else:
            assert torch.allclose(state[k], v, atol=1e-6, rtol=1e-5), k
    with torch.no_grad():
        ev = O.forward(state, g["x"], c["spec"], c["preact"], c["use_proj"], c["dropout"], training=False)
    assert torch.allclose(ev, g["eval_logits"], atol=1e-5, rtol=1e-5)


def test_oracle_bf16_autocast_matches_reference_fixture():
    c, g = CASES["wrn_tiny"], load_case("wrn_tiny")
    state = {k: v.clone() for k, v in g["init"].items()}
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        lb = O.forward(state, g["x"], c["spec"], c["preact"], c["use_proj"], 0.0, training=True)
    assert rel_l2(lb.float(), g["bf16_train_logits"]) < 1e-6


def test_spec_grammar_widths_and_downsampling():
    L = O.parse_spec("c3,160,3,1,1 r4 r4 r4 n a ap8,1,0 fc640,10")
    assert [(l["cin"], l["cout"], l["down"]) for l in L if l["kind"] == "basic"] == [
        (160, 160, False), (160, 320, True), (320, 640, True)]
    with pytest.raises(ValueError):
        O.parse_spec("c3,16,3,1,1 zz")
    plan = O.block_plan("bottleneck", 64, True, True, True)
    assert plan["convs"] == [(64, 32, 1, 1, 0), (32, 32, 3, 2, 1), (32, 128, 1, 1, 0)]
    assert plan["norms"] == [64, 32, 32]


def test_conv_flop_count_matches_survey():
    fwd, train = O.conv_train_flops("c3,160,3,1,1 r4 r4 r4 n a ap8,1,0 fc640,10", True, True, 128, 32)
    assert abs(fwd / 1e9 - 1397.00) < 0.01 and abs(train / 1e9 - 4189.9) < 0.1
    fwd, train = O.conv_train_flops("c3,16,3,1,1 n a r3 r3 r3 ap8,1,0 fc64,10", False, False, 128, 32)
    assert abs(fwd / 1e9 - 10.38) < 0.01 and abs(train / 1e9 - 31.0) < 0.1
