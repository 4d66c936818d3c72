#!/bin/bash
# ncu --set full + source page of the two-tile attention kernel (decoder causal S=4796, and ViT 28 crops)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn2_fwd -s 2 -c 1 -f -o gpurun_out/r31_attn_dec \
  python tools/gpu_check_attn.py --case dec_4796_c --variant 2 --time > gpurun_out/r31_dec.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn2_fwd -s 2 -c 1 -f -o gpurun_out/r31_attn_vit \
  python tools/gpu_check_attn.py --case vit_28 --variant 2 --time > gpurun_out/r31_vit.log 2>&1
ls -la gpurun_out/*.ncu-rep
