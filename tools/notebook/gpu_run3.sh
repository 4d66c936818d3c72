#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/tests1.log
timeout 1500 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -s > $L 2>&1
echo "ops rc=$?" >> $L
timeout 2400 python -m pytest tests/test_gpu_model.py -m gpu -q -s >> $L 2>&1
echo "model rc=$?" >> $L
tail -80 $L
