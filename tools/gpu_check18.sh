#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 400 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
run t_script python -m pytest tests/test_script_gpu.py -q -m gpu -x
TAILN=4 run other python tools/other_configs.py
cat gpurun_out/summary.txt
