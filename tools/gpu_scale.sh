#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
nvidia-smi -L | wc -l
for n in 8; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/bench_n$n.log 2>&1
  echo "bench_n$n exit=$?" | tee -a gpurun_out/summary.txt
  tail -1 gpurun_out/bench_n$n.log | cut -c1-260
done
