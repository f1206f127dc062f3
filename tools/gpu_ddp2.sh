#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 240 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-2} gpurun_out/$name.log | cut -c1-600; }
run t_ddp python -m pytest tests/test_ddp_gpu.py -q -m gpu -x
B200_WGRAD_SLAB=64 run t_wgrad64 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv_wgrad and tc"
run bench2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5
cat gpurun_out/summary.txt
