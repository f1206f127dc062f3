#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_model_gpu.py -q -m gpu -k "ragged or gradscaler or odd_batch" > gpurun_out/t_new.log 2>&1; echo "exit=$?"; tail -15 gpurun_out/t_new.log | cut -c1-300
