#!/bin/bash
# round-2 GPU call 1: new baseline-shape parity tests first (fail fast), then the whole -m gpu suite, then bench
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_baseline_shapes_gpu.py -m gpu -x -q -s > gpurun_out/r2_t_baseline.log 2>&1
echo "baseline rc=$?" >> gpurun_out/r2_t_baseline.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_baseline_shapes_gpu.py > gpurun_out/r2_t_all.log 2>&1
echo "all rc=$?" >> gpurun_out/r2_t_all.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.log 2>&1
echo "bench rc=$?" >> gpurun_out/r2_bench1.log
tail -3 gpurun_out/r2_t_baseline.log gpurun_out/r2_t_all.log gpurun_out/r2_bench1.log
