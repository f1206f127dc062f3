#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-3} gpurun_out/$name.log; }
run t_conv python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv or stem"
B200_CONV_CLUSTER=4 run t_conv_c4 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv_fprop and tc or conv_dgrad and tc"
for cs in 1 2 4; do
  TAILN=7 B200_CONV_CLUSTER=$cs BENCH_TAG=_cs$cs NO_CUDNN=$([ $cs != 1 ] && echo 1) run bench_conv_cs$cs python tools/bench_conv.py
done
run bench python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline
cat gpurun_out/summary.txt
