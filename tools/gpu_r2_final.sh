#!/bin/bash
# Round-2 evidence in one 1-GPU call: tests, smoke, bench (full line), per-kernel timings, launch list, ncu --set full
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=${T:-r02h}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${T}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${T}_smoke.log
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/${T}_bench.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/${T}_bench.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.log 2>&1; echo "ref rc=$?"
BENCH_TAG=_${T} timeout 300 python tools/bench_conv.py > gpurun_out/${T}_bench_conv.log 2>&1; echo "bench_conv rc=$?"
BENCH_TAG=_${T} timeout 300 python tools/bench_ew.py > gpurun_out/${T}_bench_ew.log 2>&1; echo "bench_ew rc=$?"
ARGS="--steps 2 --warmup 3 --sustained 0 --no-cpu-baseline --no-gpu-reference"
timeout 200 python bench.py $ARGS > gpurun_out/${T}_plain_bench.log 2>&1 &&
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 800 --csv \
   --log-file gpurun_out/${T}_launches.csv python bench.py $ARGS > gpurun_out/${T}_ncu_launches.log 2>&1
echo "launches rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1100 -c 800 --csv \
   --log-file gpurun_out/${T}_launches_warm.csv python bench.py $ARGS > gpurun_out/${T}_ncu_launches_warm.log 2>&1
echo "warm launches rc=$?"
timeout 100 python tools/kernels_once.py 1 > gpurun_out/${T}_plain_once.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2h|wgrad_tc2h|conv_tc2_|bn_act" -c 30 \
   -o gpurun_out/${T}_prof python tools/kernels_once.py 1 > gpurun_out/${T}_ncu_full.log 2>&1
echo "full rc=$?"
