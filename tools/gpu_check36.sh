#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
TAILN=1 run bench python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline
TAILN=1 B200_WGRAD_OVERLAP=0 run bench_noovl python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline
cat gpurun_out/summary.txt
