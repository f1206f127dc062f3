#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-3} gpurun_out/$name.log; }
run t_kern python -m pytest tests/test_kernels_gpu.py -q -m gpu -x
run t_model python -m pytest tests/test_model_gpu.py -q -m gpu -x
run bench python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager > gpurun_out/plain_short.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 960 -c 640 --csv \
   --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eager > gpurun_out/ncu_launches.log 2>&1
echo "ncu_launches exit=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/summary.txt
