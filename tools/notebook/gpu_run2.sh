#!/bin/bash
# attention bring-up: each variant in its own process (a trap poisons the CUDA context)
mkdir -p gpurun_out
L=gpurun_out/attn1.log
: > $L
for v in 0 1; do
  for c in dec_t128 dec_t128_c vit_t128; do
    timeout 200 python tools/gpu_check_attn.py --variant $v --case $c >> $L 2>&1
    echo "v$v $c rc=$?" >> $L
  done
done
for v in 0 1; do
  timeout 600 python tools/gpu_check_attn.py --variant $v --time >> $L 2>&1
  echo "v$v all rc=$?" >> $L
done
cat $L
