#!/bin/bash
# first GPU bring-up: GEMM correctness (small, then all shapes with timing)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gemm1.log 2>&1
timeout 300 python tools/gpu_check_gemm.py --case tile1 --case tile1_k4 --case tile1_k9 --case tile1_bn128 >> gpurun_out/gemm1.log 2>&1
echo "small rc=$?" >> gpurun_out/gemm1.log
timeout 900 python tools/gpu_check_gemm.py --time >> gpurun_out/gemm1.log 2>&1
echo "all rc=$?" >> gpurun_out/gemm1.log
cat gpurun_out/gemm1.log
