#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run22.log
: > $L
timeout 150 python tools/gpu_check_attn.py --variant 2 --case dec_t128 --case dec_t128_c --case dec_t300_c --case vit_t676 --case dec_t256 --case vit_t128 >> $L 2>&1
echo "v2 rc=$?" >> $L
timeout 200 python tools/gpu_check_attn.py --variant 2 --time --case dec_2937_c --case dec_4796_c --case vit_28 >> $L 2>&1
echo "attn rc=$?" >> $L
cat $L | cut -c1-200 | tail -14
