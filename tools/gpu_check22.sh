#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do timeout 300 python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench$i.log 2>&1; tail -1 gpurun_out/bench$i.log | cut -c1-200; done
