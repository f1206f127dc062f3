#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-3} gpurun_out/$name.log; }
run t_conv python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv or stem"
B200_WGRAD_CLUSTER=4 run t_wgrad_c4 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv_wgrad and tc"
for cs in 1 2 4; do
TAILN=1 NO_CUDNN=1 B200_WGRAD_CLUSTER=$cs BENCH_TAG=_w$cs run bench_conv_w$cs python tools/bench_conv.py
done
run t_model python -m pytest tests/test_model_gpu.py -q -m gpu -x
run bench python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline
B200_WGRAD_CLUSTER=4 run bench_w4 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline
cat gpurun_out/summary.txt
