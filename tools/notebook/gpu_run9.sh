#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run9.log
: > $L
timeout 1200 python -m pytest tests/test_gpu_idefics2.py -m gpu -x -q -s >> $L 2>&1
echo "idefics2 rc=$?" >> $L
tail -40 $L
