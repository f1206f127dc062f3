#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for d in 0 1 2; do echo "== bn_stats dbg $d"; B200_DBG_BN=$d BENCH_EW_CASES=bn_stats BENCH_TAG=_d$d timeout 100 python tools/bench_ew.py 2>&1 | grep -v Warn; done
for d in 0 1 2; do echo "== bwd dbg $d (1 = reduce only, 2 = apply only)"; B200_DBG_BWD=$d BENCH_EW_CASES=bn_act_bwd_mask_drop,bn_act_bwd_mask_add BENCH_TAG=_b$d timeout 100 python tools/bench_ew.py 2>&1 | grep -v Warn; done
