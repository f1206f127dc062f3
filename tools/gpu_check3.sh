#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -4 gpurun_out/$name.log; }
run bench python bench.py --gpus 1 --steps 30 --warmup 5
run bench_ref python bench.py --impl reference --steps 3 --warmup 1
# launch list of a short run (warm-up launches skipped), only after the plain run above exited 0
if grep -q "bench exit=0" gpurun_out/summary.txt; then
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_short.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv \
     --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
  echo "ncu_launches exit=$?" | tee -a gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
