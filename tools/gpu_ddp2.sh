#!/bin/bash
# 2-GPU validation: NCCL DDP tests + the N=2 bench line (tight timeouts: a hang must not eat the budget)
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 200 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-2} gpurun_out/$name.log | cut -c1-700; }
run t_ddp python -m pytest tests/test_ddp_gpu.py -q -m gpu -x
run bench2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5
cat gpurun_out/summary.txt
