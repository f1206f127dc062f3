#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-300; }
run t_conv python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv_fprop or conv_dgrad"
export NO_CUDNN=1
TAILN=1 BENCH_TAG=_halo run bc_halo python tools/bench_conv.py
TAILN=1 B200_CONV_HALO=0 BENCH_TAG=_nohalo run bc_nohalo python tools/bench_conv.py
TAILN=1 B200_HALO_TPB=9 BENCH_TAG=_halo9 run bc_halo9 python tools/bench_conv.py
cat gpurun_out/summary.txt
