#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-2} gpurun_out/$name.log | cut -c1-300; }
run t_kern python -m pytest tests/test_kernels_gpu.py -q -m gpu -x
B200_WGRAD_SLAB=64 run t_wgrad64 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "conv_wgrad and tc"
run t_model python -m pytest tests/test_model_gpu.py -q -m gpu -x
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench python bench.py --gpus 1 --steps 30 --warmup 5
cat gpurun_out/summary.txt
