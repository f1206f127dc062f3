#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-400} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
TAILN=25 run t_stats python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "fused or unconsumed or conv_fprop"
TAILN=8 run t_model python -m pytest tests/test_model_gpu.py -q -m gpu -x
TAILN=1 run bench python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline
TAILN=1 B200_FUSED_BN_STATS=0 run bench_nofuse python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline
cat gpurun_out/summary.txt
