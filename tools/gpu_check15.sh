#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 200 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; }
export NO_CUDNN=1
B200_WGRAD_SLAB=32 BENCH_TAG=_s3 run bc_s3 python tools/bench_conv.py
B200_WGRAD_SLAB=32 B200_WGRAD_STAGES=2 BENCH_TAG=_s2 run bc_s2 python tools/bench_conv.py
B200_WGRAD_SLAB=32 B200_WGRAD_PX=64 BENCH_TAG=_px64 run bc_px64 python tools/bench_conv.py
B200_WGRAD_SLAB=32 B200_WGRAD_PX=64 B200_WGRAD_STAGES=3 BENCH_TAG=_px64s3 run bc_px64s3 python tools/bench_conv.py
