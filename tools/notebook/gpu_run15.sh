#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run15.log
: > $L
for v in 6 10; do
  timeout 200 python tools/gpu_check_attn.py --variant $v --case dec_t300_c --case vit_t676 >> $L 2>&1
  echo "v$v rc=$?" >> $L
done
timeout 600 python tools/gpu_check_attn.py --variant 2 --variant 6 --variant 10 --time --case dec_2937_c --case dec_4796_c --case vit_28 --case vit_t676 >> $L 2>&1
echo "attn rc=$?" >> $L
timeout 600 python tools/gpu_check_gemm.py --time --case d_down --case d_down_256 --case v_fc2 --case d_o >> $L 2>&1
echo "gemm rc=$?" >> $L
cat $L | tail -30
