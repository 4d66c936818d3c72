#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run10.log
: > $L
timeout 600 python -m pytest tests/test_gpu_idefics2.py -m gpu -x -q -s >> $L 2>&1
echo "idefics2 rc=$?" >> $L
for c in ragged dec_o vit_qkv; do
  timeout 120 python tools/gpu_check_gemm.py --pair --case $c >> $L 2>&1
  echo "pair $c rc=$?" >> $L
done
timeout 900 python tools/gpu_check_gemm.py --pair --time >> $L 2>&1
echo "pair all rc=$?" >> $L
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "norm or gemm" >> $L 2>&1
echo "ops rc=$?" >> $L
timeout 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench5.json 2>> $L
tail -5 $L
