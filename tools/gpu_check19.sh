#!/bin/bash
mkdir -p gpurun_out
python tools/other_configs.py wrn-50-2-like-imagenet > gpurun_out/plain_imnet.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 600 --csv \
   --log-file gpurun_out/launches_imnet.csv python tools/other_configs.py wrn-50-2-like-imagenet > gpurun_out/ncu_imnet.log 2>&1
echo "exit=$?"
