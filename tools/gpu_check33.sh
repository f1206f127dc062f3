#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-400} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
TAILN=15 run t_kern python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "bn or dropout or subsample"
echo "== default"; timeout 120 python tools/bench_ew.py 2>&1 | grep "us "
export BENCH_EW_CASES=bn_stats,bn_act_bwd
echo "== stats bps 3 / reduce 2"; B200_BN_STATS_BPS=3 B200_BN_REDUCE_BPS=2 timeout 120 python tools/bench_ew.py 2>&1 | grep "us "
unset BENCH_EW_CASES
run t_model python -m pytest tests/test_model_gpu.py -q -m gpu -x
TAILN=1 run bench python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline
cat gpurun_out/summary.txt
