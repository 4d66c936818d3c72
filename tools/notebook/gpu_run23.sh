#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run23.log
: > $L
for c in dec_t128 dec_t128_c vit_t128 dec_t300_c vit_t676 dec_t256; do
  timeout 100 python tools/gpu_check_attn.py --variant 18 --case $c >> $L 2>&1
  echo "v18 $c rc=$?" >> $L
done
timeout 200 python tools/gpu_check_attn.py --variant 18 --variant 2 --time --case dec_2937_c --case dec_4796_c --case vit_28 >> $L 2>&1
echo "attn rc=$?" >> $L
cat $L | cut -c1-200 | tail -20
