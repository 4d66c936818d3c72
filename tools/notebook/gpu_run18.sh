#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run18.log
: > $L
timeout 900 python -m pytest tests/test_gpu_hf.py tests/test_gpu_idefics2.py tests/test_gpu_ops.py -m gpu -x -q >> $L 2>&1
echo "tests rc=$?" >> $L
timeout 600 python tools/gpu_decode_speed.py >> $L 2>&1
echo "decode rc=$?" >> $L
tail -12 $L
