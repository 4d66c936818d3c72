#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/attn2.log
: > $L
for c in dec_t128 dec_t128_c vit_t128 dec_t256 dec_t300_c; do
  timeout 200 python tools/gpu_check_attn.py --variant 2 --case $c >> $L 2>&1
  echo "v2 $c rc=$?" >> $L
done
timeout 600 python tools/gpu_check_attn.py --variant 2 --variant 1 --time >> $L 2>&1
echo "all rc=$?" >> $L
cat $L
