#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for d in 0 1 2; do
echo "== NOFIN=$d (1: no ticket/drain, 2: also no global atomics)"
B200_DBG_NOFIN=$d BENCH_CONV_NSHAPES=2 NO_CUDNN=1 BENCH_CONV_CASES=fprop,fprop_stats,fprop_res_stats BENCH_TAG=_e4 timeout 200 python tools/bench_conv.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('x'.join(map(str,d['shape'])), ' '.join(f'{k[:-3]}:{d[k]*1e3:.1f}' for k in d if k.endswith('_ms')))"
done
