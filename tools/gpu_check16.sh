#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 240 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -2 gpurun_out/$name.log | cut -c1-200; }
run t_wgrad python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv_wgrad or stem"
export NO_CUDNN=1
BENCH_TAG=_p0 run bc_p0 python tools/bench_conv.py
B200_WGRAD_SLAB=32 BENCH_TAG=_p32 run bc_p32 python tools/bench_conv.py
B200_WGRAD_MT=2 B200_WGRAD_SLAB=32 BENCH_TAG=_pmt2 run bc_pmt2 python tools/bench_conv.py
B200_WGRAD_MT=2 B200_WGRAD_PX=128 B200_WGRAD_SLAB=32 BENCH_TAG=_pmt2px128 run bc_pmt2px128 python tools/bench_conv.py
