#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-4} gpurun_out/$name.log; }
run t_kern python -m pytest tests/test_kernels_gpu.py -q -m gpu -x
run t_model python -m pytest tests/test_model_gpu.py -q -m gpu -x
TAILN=8 run bench_conv python tools/bench_conv.py
run bench python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline
cat gpurun_out/summary.txt
