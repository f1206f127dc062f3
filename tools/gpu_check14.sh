#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-3} gpurun_out/$name.log; }
run t_wgrad python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv_wgrad or stem"
TAILN=1 NO_CUDNN=1 BENCH_TAG=_mt2 run bench_conv_mt2 python tools/bench_conv.py
TAILN=1 NO_CUDNN=1 B200_WGRAD_MT=1 BENCH_TAG=_mt1 run bench_conv_mt1 python tools/bench_conv.py
TAILN=1 NO_CUDNN=1 B200_WGRAD_SLAB=32 BENCH_TAG=_mt2s32 run bench_conv_mt2s32 python tools/bench_conv.py
run t_model python -m pytest tests/test_model_gpu.py -q -m gpu -x
run bench python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline
cat gpurun_out/summary.txt
