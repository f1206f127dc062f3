#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 400 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-2} gpurun_out/$name.log | cut -c1-240; }
run t_kern python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "stem"
run t_model python -m pytest tests/test_model_gpu.py -q -m gpu -x
run bench_ov python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline
B200_WGRAD_OVERLAP=0 run bench_noov python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline
cat gpurun_out/summary.txt
