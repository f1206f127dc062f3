"""Per-shape timing of the conv kernels (CUDA events, rotating buffers larger than L2) with the
torch/cuDNN bf16 channels_last time of the same op beside it. Writes gpurun_out/bench_conv.json."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_ddp_resnet_b200 import ops, _lib  # noqa: E402

SHAPES = [  # N, H, W, C, K, R, stride, pad  (WRN-28-10, batch 128)
    (128, 32, 32, 160, 160, 3, 1, 1),
    (128, 16, 16, 320, 320, 3, 1, 1),
    (128, 8, 8, 640, 640, 3, 1, 1),
    (128, 32, 32, 160, 320, 3, 2, 1),
    (128, 16, 16, 320, 640, 3, 2, 1),
    (128, 16, 16, 160, 320, 1, 1, 0),
]

IMAGENET_SHAPES = [  # WRN-50-2-like bottleneck stacks at 224 x 224, batch 256 (bench.py --config wrn50-imagenet)
    (256, 56, 56, 512, 128, 1, 1, 0), (256, 56, 56, 128, 128, 3, 1, 1), (256, 56, 56, 128, 512, 1, 1, 0),
    (256, 28, 28, 1024, 256, 1, 1, 0), (256, 28, 28, 256, 256, 3, 1, 1), (256, 28, 28, 256, 1024, 1, 1, 0),
    (256, 14, 14, 2048, 512, 1, 1, 0), (256, 14, 14, 512, 512, 3, 1, 1), (256, 14, 14, 512, 2048, 1, 1, 0),
    (256, 7, 7, 4096, 1024, 1, 1, 0), (256, 7, 7, 1024, 1024, 3, 1, 1), (256, 7, 7, 1024, 4096, 1, 1, 0),
]
if os.environ.get("BENCH_CONV_SET") == "imagenet":
    SHAPES = IMAGENET_SHAPES


def timeit(fn, iters=24, warm=3):
    """CUDA-event time per launch with the launches captured in a CUDA graph (no host overhead)."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn(0)
        with torch.cuda.graph(g, stream=st):
            for i in range(iters):
                fn(i)
    torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    res = []
    NB = 6  # rotating buffers
    for (N, H, W, C, K, R, st, pad) in SHAPES[:int(os.environ.get("BENCH_CONV_NSHAPES", len(SHAPES)))]:
        N = int(os.environ.get("BENCH_CONV_N", N))
        P = (H + 2 * pad - R) // st + 1
        xs = [torch.randn(N, H, W, C, device="cuda").bfloat16() for _ in range(NB)]
        dys = [torch.randn(N, P, P, K, device="cuda").bfloat16() for _ in range(NB)]
        w = (torch.randn(K, R, R, C, device="cuda") * 0.02).bfloat16()
        wt = w.permute(3, 1, 2, 0).contiguous()
        flops = 2.0 * N * P * P * K * C * R * R
        row = {"shape": [N, H, W, C, K, R, st, pad], "gflop": flops / 1e9}
        mean, invstd = ops.bn_stats(xs[0], 1e-5)
        masks = [ops.bn_act_fwd(x, mean, invstd, torch.ones(C, device="cuda"), torch.zeros(C, device="cuda"),
                                relu=True, dropout_p=0.3, seed=1, want_mask=True)[1] for x in xs]
        for name, fn in [
            ("fprop", lambda i: ops.conv_fprop(xs[i % NB], w, st, pad, algo=_lib.ALGO_TC)),
            ("fprop_stats", lambda i: ops.conv_fprop(xs[i % NB], w, st, pad, algo=_lib.ALGO_TC, want_stats=True)),
            ("fprop_res_stats", lambda i: ops.conv_fprop(xs[i % NB], w, st, pad, algo=_lib.ALGO_TC, want_stats=True,
                                                         residual=dys[(i + 1) % NB])),
            ("dgrad", lambda i: ops.conv_dgrad(dys[i % NB], wt, (H, W), st, pad, algo=_lib.ALGO_TC)),
            ("dgrad_bnbwd", lambda i: ops.conv_dgrad_bn_bwd(dys[i % NB], wt, (H, W), st, pad, algo=_lib.ALGO_TC,
                                                            x_bn=xs[(i + 1) % NB], mask=masks[(i + 1) % NB],
                                                            mean=mean, invstd=invstd, dropout_p=0.3)),
            ("wgrad", lambda i: ops.conv_wgrad(dys[i % NB], xs[i % NB], R, R, st, pad, algo=_lib.ALGO_TC)),
        ]:
            only = os.environ.get("BENCH_CONV_CASES")
            if only and name not in only.split(","):
                continue
            ms = timeit(fn)
            row[name + "_ms"] = ms
            row[name + "_tflops"] = flops / ms / 1e9
        if os.environ.get("NO_CUDNN"):
            print(json.dumps(row), flush=True)
            res.append(row)
            continue
        # cuDNN bf16 channels_last for context
        xc = [x.permute(0, 3, 1, 2) for x in xs]
        wc = w.permute(0, 3, 1, 2)
        dyc = [d.permute(0, 3, 1, 2) for d in dys]
        row["cudnn_fprop_ms"] = timeit(lambda i: F.conv2d(xc[i % NB], wc, stride=st, padding=pad))
        row["cudnn_dgrad_ms"] = timeit(lambda i: torch.nn.grad.conv2d_input(
            (N, C, H, W), wc, dyc[i % NB], stride=st, padding=pad))
        row["cudnn_wgrad_ms"] = timeit(lambda i: torch.nn.grad.conv2d_weight(
            xc[i % NB], (K, C, R, R), dyc[i % NB], stride=st, padding=pad))
        print(json.dumps(row), flush=True)
        res.append(row)
    os.makedirs("gpurun_out", exist_ok=True)
    tag = os.environ.get("BENCH_TAG", "")
    with open(f"gpurun_out/bench_conv{tag}.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
