#!/bin/bash
# BN forward variants: kernel tests, per-kernel timings and whole-step timings per variant
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -x -q > gpurun_out/bn_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/bn_tests.log
ARGS="--steps 30 --warmup 5 --sustained 0 --no-cpu-baseline --no-gpu-reference"
for v in 0 1 2 3; do
  echo "== variant $v"
  B200_BN_DROP_VARIANT=$v BENCH_EW_CASES=bn_act_fwd,bn_act_fwd_dropout,bn_act_fwd_drop_mask BENCH_TAG=_v$v timeout 200 python tools/bench_ew.py 2>&1 | grep -v Warn
  B200_BN_DROP_VARIANT=$v timeout 300 python bench.py $ARGS 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('step ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['clocks']['sm_mhz'])"
done
