#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
cp pytorch_ddp_resnet_b200/libb200resnet.so variants/lib_v0.so
ARGS="--steps 30 --warmup 5 --sustained 0 --no-cpu-baseline --no-gpu-reference"
for v in 0 1 2 3 4; do
  echo "== ldg256 variant $v"
  cp variants/lib_v$v.so pytorch_ddp_resnet_b200/libb200resnet.so
  NO_CUDNN=1 BENCH_CONV_CASES=fprop,fprop_res_stats,dgrad_bnbwd BENCH_TAG=_ldg$v timeout 200 python tools/bench_conv.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('x'.join(map(str,d['shape'])), ' '.join(f'{k[:-3]}:{d[k]*1e3:.1f}' for k in d if k.endswith('_ms')))"
  timeout 300 python bench.py $ARGS 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('step ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['clocks']['sm_mhz'])"
done
cp variants/lib_v0.so pytorch_ddp_resnet_b200/libb200resnet.so
