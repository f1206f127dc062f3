#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for n in 64 128 256; do
echo "== N=$n"
BENCH_CONV_N=$n BENCH_CONV_NSHAPES=3 NO_CUDNN=1 BENCH_CONV_CASES=fprop,fprop_stats,fprop_res_stats,dgrad,dgrad_bnbwd BENCH_TAG=_e3 timeout 200 python tools/bench_conv.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('x'.join(map(str,d['shape'])), ' '.join(f'{k[:-3]}:{d[k]*1e3:.1f}' for k in d if k.endswith('_ms')))"
done
