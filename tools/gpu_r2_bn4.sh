#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_baseline_shapes_gpu.py -m gpu -x -q > gpurun_out/bn4_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/bn4_tests.log
ARGS="--steps 30 --warmup 5 --sustained 0 --no-cpu-baseline --no-gpu-reference"
BENCH_TAG=_bn4 timeout 200 python tools/bench_ew.py 2>&1 | grep -v Warn
for d in 1 2; do echo "== bwd dbg $d (1 = reduce only, 2 = apply only)"; B200_DBG_BWD=$d BENCH_EW_CASES=bn_act_bwd_mask_drop,bn_act_bwd_mask_add BENCH_TAG=_b$d timeout 100 python tools/bench_ew.py 2>&1 | grep -v Warn; done
for i in 1 2 3; do timeout 300 python bench.py $ARGS 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('step ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['clocks']['sm_mhz'])"; done
