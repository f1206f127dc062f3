#!/bin/bash
# First GPU bring-up: kernel tests in separate processes + conv micro-benchmark.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -5 gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run t_misc   python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not conv_fprop and not conv_dgrad and not conv_wgrad"
run t_direct python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "direct and (conv_fprop or conv_dgrad or conv_wgrad)"
run t_tc_fprop python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv_fprop and tc"
run t_tc_dgrad python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv_dgrad and tc"
run t_tc_wgrad python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv_wgrad and tc"
run bench_conv python tools/bench_conv.py
cat gpurun_out/summary.txt
