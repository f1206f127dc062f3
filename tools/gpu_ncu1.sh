#!/bin/bash
mkdir -p gpurun_out
python tools/kernels_once.py 2 > gpurun_out/plain_kernels_once.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|wgrad_tc|bn_stats_partial|bn_act" -s 27 -c 27 \
   -o gpurun_out/prof_r1_kernels python tools/kernels_once.py 2 > gpurun_out/ncu_full.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
