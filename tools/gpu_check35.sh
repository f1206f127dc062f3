#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-300} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
TAILN=12 run t_conv python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv"
TAILN=12 B200_HALO_MT=1 run t_conv_mt1 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv_fprop or conv_dgrad"
export NO_CUDNN=1
TAILN=1 BENCH_TAG=_mt2 run bc_mt2 python tools/bench_conv.py
TAILN=1 B200_HALO_MT=1 BENCH_TAG=_mt1 run bc_mt1 python tools/bench_conv.py
TAILN=1 B200_HALO_MT=1 B200_HALO_PW=16 BENCH_TAG=_mt1pw16 run bc_mt1pw16 python tools/bench_conv.py
unset NO_CUDNN
run t_model python -m pytest tests/test_model_gpu.py -q -m gpu -x
TAILN=1 run bench python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline
cat gpurun_out/summary.txt
