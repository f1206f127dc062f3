#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "fused_bn_backward" > gpurun_out/e_tests1.log 2>&1; echo "new test rc=$?"; tail -15 gpurun_out/e_tests1.log
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_baseline_shapes_gpu.py -m gpu -x -q > gpurun_out/e_tests2.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/e_tests2.log
ARGS="--steps 30 --warmup 5 --sustained 0 --no-cpu-baseline --no-gpu-reference"
for f in 1 0 1 0; do
  echo "== fused bn bwd $f"
  B200_FUSED_BN_BWD=$f timeout 300 python bench.py $ARGS 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('step ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['clocks']['sm_mhz'], 'launches', d['gpu_launches'], 'loss', d['e2e']['last_loss'])"
done
