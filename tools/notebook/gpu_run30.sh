#!/bin/bash
# attention softmax turn protocol: correctness + timing, off (variant 34) vs on (variant 18)
mkdir -p gpurun_out
timeout 300 python tools/gpu_check_attn.py --variant 34 --variant 18 --time > gpurun_out/r30_attn.log 2>&1
echo "rc=$?"; cat gpurun_out/r30_attn.log | cut -c1-200
