#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run11.log
: > $L
C="--case d_qkv --case d_qkv_192 --case d_o --case d_o_256 --case d_gateup --case d_down --case d_down_256 --case v_qkv --case v_out --case v_out_256 --case v_fc1 --case v_fc2 --case v_fc2_256 --case p_1 --case ragged --case sw192 --case r224"
echo "== single" >> $L
timeout 600 python tools/gpu_check_gemm.py --time $C >> $L 2>&1
echo "== pair" >> $L
timeout 600 python tools/gpu_check_gemm.py --pair --time $C >> $L 2>&1
timeout 600 python bench.py --steps 5 --no-cpu-baseline --gemm-pair 0 > gpurun_out/bench6_single.json 2>> $L
timeout 600 python bench.py --steps 5 --no-cpu-baseline --gemm-pair 1 > gpurun_out/bench6_pair.json 2>> $L
tail -3 $L
