#!/bin/bash
mkdir -p gpurun_out
export B200_CONV_CLUSTER=1
python tools/conv_once.py > gpurun_out/plain_conv_once.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|wgrad_tc" -s 4 -c 4 \
   -o gpurun_out/prof_r1_conv2 python tools/conv_once.py > gpurun_out/ncu_conv2.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/ncu_conv2.log
