#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -6 gpurun_out/$name.log; }
run t_graph python -m pytest tests/test_model_gpu.py -q -m gpu -k "graph or dropout" -x
run bench python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline
run bench_eager python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --eager
cat gpurun_out/summary.txt
