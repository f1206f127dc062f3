#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-600} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" | tee -a gpurun_out/summary.txt; tail -${TAILN:-3} gpurun_out/$name.log | cut -c1-400; }
TAILN=6 run t_full python -m pytest tests -q -m gpu -x
TAILN=1 run bench python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline
TAILN=2 run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
cat gpurun_out/summary.txt
