#!/bin/bash
# quick GPU check: the -m gpu suite (fail fast), the elementwise micro-benchmark, a short bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-q}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_${TAG}_tests.log
tail -n 4 gpurun_out/r2_${TAG}_tests.log
BENCH_TAG=_${TAG} timeout 300 python tools/bench_ew.py > gpurun_out/r2_${TAG}_ew.log 2>&1
tail -n 20 gpurun_out/r2_${TAG}_ew.log
timeout 300 python bench.py --steps 30 --warmup 5 --sustained 100 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2_${TAG}_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/r2_${TAG}_bench.log
tail -n 2 gpurun_out/r2_${TAG}_bench.log | cut -c1-700
