#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -4 gpurun_out/$name.log; }
run t_kern python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "stem or bn_ or tc"
run t_model python -m pytest tests/test_model_gpu.py -q -m gpu
run stepprof python tools/step_profile.py
run bench python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline
cat gpurun_out/summary.txt
