#!/bin/bash
# 8-GPU: exposed-communication sweep over NCCL CTA budgets
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for ctas in 8 4 16; do
  export NCCL_MAX_CTAS=$ctas
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2971$((RANDOM % 10)) \
      bench.py --gpus 8 --steps 30 --warmup 5 --sustained 0 > gpurun_out/r2_n8_$ctas.log 2>&1
  echo "ctas=$ctas rc=$?"
  grep '^{' gpurun_out/r2_n8_$ctas.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:d.get(k) for k in ('ms_per_step','exposed_comm_ms','ms_per_step_without_collectives','value')})"
done
