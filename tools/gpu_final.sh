#!/bin/bash
# Round-end validation + measurements + profiles in one call (1 GPU).
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-600; }
TAILN=4 run t_full python -m pytest tests -q -m gpu -x
TAILN=2 run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TAILN=1 run bench_final python bench.py --gpus 1
NO_CUDNN=1 TAILN=1 BENCH_TAG=_final run bc_final python tools/bench_conv.py
TAILN=20 BENCH_TAG=_final run ew_final python tools/bench_ew.py
# profiles: launch list of the bench command, then --set full of the top kernels
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench_short.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 1000 --csv \
   --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "launches exit=$?" | tee -a gpurun_out/summary.txt
python tools/kernels_once.py 1 > gpurun_out/plain_kernels_once.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|wgrad_tc|bn_" -c 12 \
   -o gpurun_out/prof_r01_final python tools/kernels_once.py 1 > gpurun_out/ncu_final.log 2>&1
echo "full exit=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/summary.txt
