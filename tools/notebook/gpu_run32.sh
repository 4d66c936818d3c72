#!/bin/bash
# attention v3 (64-key S halves, issuer per tile): correctness + timing vs v2
mkdir -p gpurun_out
timeout 300 python tools/gpu_check_attn.py --variant 64 --variant 2 --time > gpurun_out/r32_attn.log 2>&1
echo "rc=$?"; cat gpurun_out/r32_attn.log | cut -c1-220
