#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run17.log
: > $L
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q >> $L 2>&1
echo "fullsize rc=$?" >> $L
timeout 600 python tools/gpu_decode_speed.py >> $L 2>&1
echo "decode rc=$?" >> $L
tail -25 $L
