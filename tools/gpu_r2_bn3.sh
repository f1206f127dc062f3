#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_baseline_shapes_gpu.py -m gpu -x -q > gpurun_out/bn3_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/bn3_tests.log
ARGS="--steps 30 --warmup 5 --sustained 0 --no-cpu-baseline --no-gpu-reference"
for b in 1 0; do
  echo "== balanced $b"
  B200_BN_BALANCED=$b BENCH_TAG=_bal$b timeout 200 python tools/bench_ew.py 2>&1 | grep -v Warn | grep -v "bn_act_bwd \|bwd_dropout\|bwd_addend\|fwd_dropout "
  for i in 1 2; do B200_BN_BALANCED=$b timeout 300 python bench.py $ARGS 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('step ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['clocks']['sm_mhz'])"; done
done
