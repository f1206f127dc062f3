"""
CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares (no compute calls), spec parsing / state_dict layout against the golden fixtures made from
the reference, config and checkpoint glue, and the N > 1 host logic under gloo with world_size 2.
"""
import os
import re
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.golden_util import CASES, load_case  # noqa: E402


def test_library_exports_every_header_symbol():
    from pytorch_ddp_resnet_b200 import _lib
    header = open(os.path.join(ROOT, "include", "b200resnet.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in the header"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.b200_version() >= 100
    assert _lib.launch_count() >= 0
    # pure host-side queries are callable without a GPU
    assert lib.b200_conv2d_tc_supported(0, 128, 32, 32, 160, 160, 3, 3, 1, 1) == 1
    assert lib.b200_conv2d_tc_supported(0, 128, 32, 32, 3, 160, 3, 3, 1, 1) == 0
    # stride-2 fprop / wgrad read x through strided TMA boxes (no workspace); only the multi-launch dgrad
    # fallback needs room for its four output phases
    assert lib.b200_conv2d_workspace_bytes(0, 128, 32, 32, 160, 320, 3, 3, 2, 1, 0) == 0
    assert lib.b200_conv2d_workspace_bytes(1, 128, 32, 32, 160, 320, 3, 3, 2, 1, 0) == 128 * 32 * 32 * 160 * 2
    assert lib.b200_bn_workspace_bytes(131072, 160) > 0
    # deterministic mode: the flag value matches the header, and the wgrad workspace grows by the ordered-reduction
    # partials (whole filter-gradient copies, one per pixel-range split, + the bias-gradient partials)
    flag = int(re.search(r"B200_ALGO_DETERMINISTIC\s*=\s*(0x[0-9a-fA-F]+|\d+)", header).group(1), 0)
    assert flag == _lib.ALGO_DETERMINISTIC
    for (N, H, C, K, R, st, pad) in [(128, 32, 160, 160, 3, 1, 1), (128, 8, 640, 640, 3, 1, 1),
                                     (128, 32, 160, 320, 3, 2, 1), (128, 32, 3, 160, 3, 1, 1)]:
        base = lib.b200_conv2d_workspace_bytes(2, N, H, H, C, K, R, R, st, pad, 0)
        det = lib.b200_conv2d_workspace_bytes(2, N, H, H, C, K, R, R, st, pad, flag)
        extra = det - base
        assert extra > 0 and extra < (1 << 30), (N, H, C, K, extra)
        if C >= 16:
            dw_bytes = K * R * R * C * 4
            assert extra >= 2 * dw_bytes and (extra - extra % dw_bytes) // dw_bytes <= 64, (extra, dw_bytes)
    # the flag changes nothing for fprop / dgrad workspaces
    assert lib.b200_conv2d_workspace_bytes(1, 128, 32, 32, 160, 320, 3, 3, 2, 1, flag) == 128 * 32 * 32 * 160 * 2


def test_deterministic_mode_flag_plumbing():
    """ops.deterministic() / torch.use_deterministic_algorithms / B200_DETERMINISTIC or the flag into every conv
    algo argument (host logic only: no kernel runs here)."""
    from pytorch_ddp_resnet_b200 import _lib, ops
    base = ops.conv_algo()
    assert not ops.is_deterministic() and ops._algo_flags(None) == base
    with ops.deterministic():
        assert ops.is_deterministic()
        assert ops._algo_flags(None) == base | _lib.ALGO_DETERMINISTIC
        assert ops._algo_flags(_lib.ALGO_DIRECT) == _lib.ALGO_DIRECT | _lib.ALGO_DETERMINISTIC
        with ops.deterministic(False):
            assert ops._algo_flags(_lib.ALGO_TC) == _lib.ALGO_TC
        assert ops.is_deterministic()
    assert not ops.is_deterministic()
    torch.use_deterministic_algorithms(True)
    try:
        assert ops.is_deterministic() and ops._algo_flags(None) & _lib.ALGO_DETERMINISTIC
    finally:
        torch.use_deterministic_algorithms(False)
    assert ops._algo_flags(None) == base


def test_no_cpu_fallback():
    from pytorch_ddp_resnet_b200._lib import B200Error
    from pytorch_ddp_resnet_b200.architectures.resnet import ResNet
    m = ResNet("c3,16,3,1,1 r1 n a ap32,1,0 fc16,10", True, True, 0.0)
    with pytest.raises(B200Error):
        m(torch.randn(2, 3, 32, 32))


@pytest.mark.parametrize("case", list(CASES))
def test_state_dict_layout_matches_reference_fixture(case):
    from pytorch_ddp_resnet_b200.architectures.resnet import ResNet
    c, g = CASES[case], load_case(case)
    m = ResNet(c["spec"], c["preact"], c["use_proj"], c["dropout"])
    sd = m.state_dict()
    assert list(sd.keys()) == list(g["init"].keys())
    for k, v in g["init"].items():
        assert tuple(sd[k].shape) == tuple(v.shape) and sd[k].dtype == v.dtype, k
    assert [n for n, _ in m.named_parameters()] == [k for k in g["grad"].keys()]
    m.load_state_dict(g["init"])
    for k, v in m.state_dict().items():
        assert torch.equal(v, g["init"][k]), k
    # conv filters stay physically KRSC (channels_last) after loading reference weights
    for mod in m.modules():
        if mod.__class__.__name__ in ("Conv2d", "ConvStem"):
            assert mod.weight.permute(0, 2, 3, 1).is_contiguous()


def test_init_distributions_follow_reference_quirk():
    """Only top-level convs get kaiming-normal; block convs keep U(+-1/sqrt(fan_in))."""
    from pytorch_ddp_resnet_b200.architectures.resnet import ResNet
    torch.manual_seed(0)
    m = ResNet("c3,160,3,1,1 r1 n a ap32,1,0 fc160,10", True, True, 0.0)
    stem, blk = m._architecture[0], m._architecture[1][0]
    assert abs(stem.weight.std().item() - (2.0 / 27) ** 0.5) < 0.02
    bound = 1.0 / (160 * 9) ** 0.5
    assert blk._conv1.weight.abs().max().item() <= bound + 1e-6
    assert abs(blk._conv1.weight.std().item() - bound / 3 ** 0.5) < 1e-3


def test_spec_grammar_errors_and_widths():
    from pytorch_ddp_resnet_b200.architectures.resnet import ResNet, tokenize
    assert tokenize("c3,16,3,1,1 n a mp3,2,1 r2 b1 ap8,1,0 fc64,10")[-1] == ("f", (64, 10))
    with pytest.raises(ValueError):
        tokenize("c3,16,3,1,1 xx")
    m = ResNet("c3,16,3,1,1 r2 r2 b1", False, True, 0.1)
    assert m._architecture[2][0]._downsample and not m._architecture[2][1]._downsample
    assert m._architecture[2][0]._out_channels == 32
    assert not m._architecture[3][0]._downsample  # a bottleneck stack after a basic stack does not


def test_config_and_checkpoint_roundtrip(tmp_path):
    from pytorch_ddp_resnet_b200.utils.config_util import ConfigParser
    from pytorch_ddp_resnet_b200.utils import checkpoint_util as C
    cfg = ConfigParser(defaults={"mode": "train"})
    cfg.read(os.path.join(ROOT, "models_dir", "wrn-28-10-dropout_cifar10", "config.yaml"))
    assert cfg.get("architecture_spec") == "c3,160,3,1,1 r4 r4 r4 n a ap8,1,0 fc640,10"
    assert dict(**cfg)["dropout_prob"] == 0.3 and len(cfg) >= 22
    with pytest.raises(KeyError):
        cfg.get("missing")
    strat = C.get_checkpoint_strategy("FrequencyCheckpointStrategy", {"unit": "batch", "frequency": 2})
    assert [strat.observe(unit="batch", loss=1.0) for _ in range(4)] == [True, False, True, False]
    lin = torch.nn.Linear(2, 2)
    for s in range(1, 8):
        C.save_checkpoints(str(tmp_path), {"classifier": lin, "checkpoint_strategy": strat, "scaler": None}, s)
    names = sorted(os.listdir(tmp_path))
    assert "classifier_7.pth" in names and "classifier_2.pth" not in names and len(names) == 10
    lin2 = torch.nn.Linear(2, 2)
    strat2 = C.get_checkpoint_strategy("FrequencyCheckpointStrategy", {"unit": "batch", "frequency": 2})
    assert C.maybe_load_checkpoints(str(tmp_path), {"classifier": lin2, "checkpoint_strategy": strat2}, "cpu", None) == 7
    assert torch.equal(lin2.weight, lin.weight) and strat2.batch_step == 4
    perf = C.get_checkpoint_strategy("PerformanceCheckpointStrategy", {"unit": "epoch"})
    assert [perf.observe(unit="epoch", loss=l) for l in (1.0, 2.0, 0.5)] == [True, False, True]


def test_lagged_metrics_fifo():
    from pytorch_ddp_resnet_b200.algos.metrics import LaggedMetrics
    lag = LaggedMetrics(world_size=1)
    for i in range(3):
        lag.push(i, {"loss": torch.tensor(float(i)), "top1_err": torch.tensor(0.5)})
        if i == 0:
            assert not lag.ready()
    assert lag.ready() and len(lag) == 3
    tags = []
    while len(lag):
        tag, m = lag.pop()
        tags.append(tag)
        assert m["loss"] == float(tag) and m["top1_err"] == 0.5
    assert tags == [0, 1, 2]


def test_plan_buckets_layout():
    from pytorch_ddp_resnet_b200.utils.graph_util import plan_buckets
    from pytorch_ddp_resnet_b200.architectures.resnet import ResNet
    m = ResNet("c3,160,3,1,1 r4 r4 r4 n a ap8,1,0 fc640,10", True, True, 0.3)
    sizes = [p.numel() for p in m.parameters()][::-1]
    offsets, owner, ranges = plan_buckets(sizes)
    assert len(offsets) == len(owner) == len(sizes) == 80
    assert ranges[0][0] == 0 and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    assert ranges[-1][1] >= sum(sizes) and all(o % 4 == 0 for o in offsets)   # 16-byte aligned slots
    assert owner == sorted(owner) and owner[-1] == len(ranges) - 1
    mb = [(hi - lo) * 4 / 2 ** 20 for lo, hi in ranges]
    assert all(20 < v < 40 for v in mb[:-1]) and mb[-1] <= 25 / 4 + 1e-6   # small exposed tail bucket
    assert plan_buckets([]) == ([], [], []) and plan_buckets([10])[2] == [(0, 12)]


# ---- N > 1 host logic on CPU: gloo, world_size 2 -------------------------------------------------
def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pytorch_ddp_resnet_b200.algos import metrics, training
    from pytorch_ddp_resnet_b200.utils import checkpoint_util as C
    from oracle import resnet_oracle as O

    # (1) packed metric all-reduce == per-metric means over ranks
    vals = {"loss": torch.tensor(1.0 + rank), "top1_err": torch.tensor(0.25 * rank), "top5_err": torch.tensor(0.0)}
    gm = metrics.global_means(vals, world)
    assert abs(gm["loss"] - 1.5) < 1e-6 and abs(gm["top1_err"] - 0.125) < 1e-6

    # (2) training_loop host logic with a CPU stand-in model (the kernels need a GPU): DDP(gloo)
    #     gradient averaging, microbatch accumulation, stepping, epoch advance on every rank,
    #     checkpoint numbering. The loss function is swapped for the oracle's (CPU) one.
    training.compute_losses_and_metrics = lambda logits, labels: O.losses_and_metrics(logits, labels)
    import pytorch_ddp_resnet_b200.algos.evaluation as ev
    ev.compute_losses_and_metrics = training.compute_losses_and_metrics
    torch.manual_seed(0)
    model = torch.nn.parallel.DistributedDataParallel(torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(12, 10)))
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(0)
    ds = torch.utils.data.TensorDataset(torch.randn(32, 3, 2, 2, generator=g), torch.randint(0, 10, (32,), generator=g))
    sampler = torch.utils.data.distributed.DistributedSampler(ds, world, rank, shuffle=True, seed=0)
    epochs_seen = []
    orig = sampler.set_epoch
    sampler.set_epoch = lambda e: (epochs_seen.append(e), orig(e))
    dl = torch.utils.data.DataLoader(ds, batch_size=4, sampler=sampler)
    strat = C.get_checkpoint_strategy("FrequencyCheckpointStrategy", {"unit": "batch", "frequency": 3})
    training.training_loop(rank=rank, world_size=world, device="cpu", sampler_train=sampler, sampler_test=sampler,
                           dl_train=dl, dl_test=dl, classifier=model, optimizer=opt, scaler=None, scheduler=None,
                           scheduler_step_unit="none", checkpoint_strategy=strat, checkpoint_dir=tmp,
                           num_microbatches=2, global_step=0, max_steps=5, log_dir=None)
    assert epochs_seen == [0, 1, 2], epochs_seen  # 2 optimizer steps per epoch (4 microbatches / 2)
    # replicas stay identical (DDP averaged the gradients)
    w = model.module[1].weight.detach().clone()
    gathered = [torch.zeros_like(w) for _ in range(world)]
    dist.all_gather(gathered, w)
    assert torch.equal(gathered[0], gathered[1])
    # (3) FlatGradReducer (the graph-mode gradient exchange): reverse-order flat buffer, bucketed
    #     all-reduce issued from post-accumulate-grad hooks, .grad ends up as the averaged slot
    from pytorch_ddp_resnet_b200.utils.graph_util import FlatGradReducer
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 50), torch.nn.ReLU(), torch.nn.Linear(50, 30), torch.nn.ReLU(),
                              torch.nn.Linear(30, 4))
    red = FlatGradReducer(net, torch.device("cpu"), world, bucket_bytes=4096)
    assert len(red.ranges) >= 3 and red.ranges[0][0] == 0
    assert red.bucket_of[id(net[4].bias)] == 0 and red.bucket_of[id(net[0].weight)] == len(red.ranges) - 1
    xg = torch.Generator().manual_seed(10 + rank)
    xin = torch.randn(5, 6, generator=xg)
    for _ in range(2):   # twice: begin() must re-arm the buckets
        for prm in net.parameters():
            prm.grad = None
        red.begin()
        net(xin).square().sum().backward()
        red.finish()
        assert all(red.launched) and red.copied > 0
        mine = [prm.grad.clone() for prm in net.parameters()]
        for prm in net.parameters():
            prm.grad = None
        net(xin).square().sum().backward()   # hooks inactive: plain local gradients
        for prm, avg in zip(net.parameters(), mine):
            loc = prm.grad.clone()
            dist.all_reduce(loc)
            assert torch.allclose(avg, loc / world, atol=1e-6), "reducer != mean of local gradients"
    red.detach()
    dist.barrier()
    if rank == 0:
        names = sorted(os.listdir(tmp))
        assert "classifier_1.pth" in names and "classifier_4.pth" in names, names
        sd = torch.load(os.path.join(tmp, "classifier_4.pth"))
        assert all(k.startswith("module.") for k in sd)  # saved through the DDP wrapper, as the reference
    dist.destroy_process_group()


def test_gloo_world_size_2_host_logic(tmp_path):
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
