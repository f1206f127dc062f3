#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-400} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
TAILN=15 run t_all python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu -x
TAILN=1 run bench_pdl python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline
TAILN=1 B200_PDL=0 run bench_nopdl python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline
echo "== ew pdl"; timeout 120 python tools/bench_ew.py 2>&1 | grep "us " | grep "bn_stats \|bn_act_fwd \|bn_act_bwd "
cat gpurun_out/summary.txt
