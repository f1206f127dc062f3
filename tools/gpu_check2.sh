#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -8 gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run t_misc   python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not conv_fprop and not conv_dgrad and not conv_wgrad"
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run t_model python -m pytest tests/test_model_gpu.py -q -m gpu -s
cat gpurun_out/summary.txt
