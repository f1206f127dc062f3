#!/bin/bash
# round-2 GPU call (2 GPUs): DDP parity tests, then bench at N=2 (with exposed-comm measurement)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ddp_gpu.py -m gpu -x -q -s > gpurun_out/r2_t_ddp.log 2>&1
echo "ddp rc=$?" >> gpurun_out/r2_t_ddp.log
tail -n 5 gpurun_out/r2_t_ddp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 2 --steps 30 --warmup 5 --sustained 100 > gpurun_out/r2_bench_n2.log 2>&1
echo "bench2 rc=$?" >> gpurun_out/r2_bench_n2.log
tail -n 3 gpurun_out/r2_bench_n2.log
