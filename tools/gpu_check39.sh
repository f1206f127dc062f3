#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
TAILN=12 run t_conv python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv"
export NO_CUDNN=1
TAILN=1 BENCH_TAG=_ew8 run bc_ew8 python tools/bench_conv.py
unset NO_CUDNN
TAILN=1 run bench python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline
TAILN=1 B200_FUSED_BN_STATS=0 run bench_nofuse python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline
cat gpurun_out/summary.txt
