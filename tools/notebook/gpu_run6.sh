#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/gemm2.log
: > $L
timeout 300 python tools/gpu_check_gemm.py --case tile1 --case r224 --case r192 --case r160 --case sw192 --case sw128 --case ragged128 >> $L 2>&1
echo "small rc=$?" >> $L
timeout 900 python tools/gpu_check_gemm.py --time >> $L 2>&1
echo "all rc=$?" >> $L
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -x -q -k "not full_depth" >> $L 2>&1
echo "tests rc=$?" >> $L
timeout 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench3.json 2>> $L
tail -5 $L
