#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_baseline_shapes_gpu.py -m gpu -x -q > gpurun_out/e2_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/e2_tests.log
NO_CUDNN=1 BENCH_CONV_CASES=fprop,fprop_stats,fprop_res_stats,dgrad_bnbwd BENCH_TAG=_e2 timeout 200 python tools/bench_conv.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('x'.join(map(str,d['shape'])), ' '.join(f'{k[:-3]}:{d[k]*1e3:.1f}' for k in d if k.endswith('_ms')))"
ARGS="--steps 30 --warmup 5 --sustained 0 --no-cpu-baseline --no-gpu-reference"
for i in 1 2 3; do timeout 300 python bench.py $ARGS 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('step ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['clocks']['sm_mhz'])"; done
