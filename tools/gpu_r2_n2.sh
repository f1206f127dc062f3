#!/bin/bash
# 2-GPU: DDP tests, then the N=2 bench with a few NCCL CTA budgets (exposed-communication sweep)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ddp_gpu.py -m gpu -x -q > gpurun_out/r2_ddp2.log 2>&1
echo "ddp rc=$?" | tee -a gpurun_out/r2_ddp2.log
tail -n 3 gpurun_out/r2_ddp2.log
for ctas in default 4 8 16; do
  if [ "$ctas" != default ]; then export NCCL_MAX_CTAS=$ctas; fi
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$((RANDOM % 10)) \
      bench.py --gpus 2 --steps 30 --warmup 5 --sustained 0 > gpurun_out/r2_n2_$ctas.log 2>&1
  echo "ctas=$ctas rc=$?"
  grep '^{' gpurun_out/r2_n2_$ctas.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:d.get(k) for k in ('ms_per_step','exposed_comm_ms','ms_per_step_without_collectives','value')})"
done
