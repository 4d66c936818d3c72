#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn3_fwd -s 2 -c 1 -f -o gpurun_out/r33_attn3_dec \
  python tools/gpu_check_attn.py --case dec_4796_c --variant 64 --time > gpurun_out/r33_dec.log 2>&1
ls -la gpurun_out/r33*.ncu-rep
