#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-240} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-600; }
TAILN=12 run t_wgrad python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv_wgrad"
TAILN=12 B200_WGRAD_PW=16 run t_wgrad16 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv_wgrad"
export NO_CUDNN=1
TAILN=1 BENCH_TAG=_wh10 run bc_wh10 python tools/bench_conv.py
TAILN=1 B200_WGRAD_PW=16 BENCH_TAG=_wh16 run bc_wh16 python tools/bench_conv.py
TAILN=1 B200_WGRAD_HALO=0 BENCH_TAG=_wh0 run bc_wh0 python tools/bench_conv.py
cat gpurun_out/summary.txt
