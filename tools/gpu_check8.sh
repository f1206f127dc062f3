#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 240 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-3} gpurun_out/$name.log; }
run t_conv python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv or stem"
TAILN=7 NO_CUDNN=1 BENCH_TAG=_pair run bench_conv_pair python tools/bench_conv.py
run bench python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline
cat gpurun_out/summary.txt
