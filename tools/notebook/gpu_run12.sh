#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run12.log
: > $L
C="--case v_out --case v_out_b --case v_out_n --case v_out_n256 --case v_k2304_n --case d_qkv --case d_o --case v_patch"
echo "== pair" >> $L
timeout 600 python tools/gpu_check_gemm.py --pair --time $C >> $L 2>&1
echo "== single" >> $L
timeout 600 python tools/gpu_check_gemm.py --single --time $C >> $L 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -k "not full_depth" >> $L 2>&1
echo "tests rc=$?" >> $L
timeout 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench7.json 2>> $L
tail -3 $L
