#!/bin/bash
mkdir -p gpurun_out
export BENCH_EW_CASES=bn_act_fwd,bn_act_fwd_dropout
for cfg in 2,3,12 4,4,12 4,4,4 4,4,8 4,3,12 8,2,2 8,2,8; do
  echo "== B200_EW_FWD=$cfg"
  B200_EW_FWD=$cfg timeout 120 python tools/bench_ew.py 2>&1 | grep "us "
done
export BENCH_EW_CASES=bn_stats,bn_act_bwd
for bps in 2 4 6; do
  echo "== B200_BN_BPS=$bps"
  B200_BN_BPS=$bps timeout 120 python tools/bench_ew.py 2>&1 | grep "us "
done
