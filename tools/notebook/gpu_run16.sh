#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run16.log
: > $L
timeout 300 python tools/gpu_check_attn.py --variant 2 --case dec_t128_c --case dec_t300_c --case vit_t676 --case dec_t256 >> $L 2>&1
echo "v2 rc=$?" >> $L
timeout 600 python tools/gpu_check_attn.py --variant 2 --variant 6 --variant 10 --time --case dec_2937_c --case dec_4796_c --case vit_28 >> $L 2>&1
echo "attn rc=$?" >> $L
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_hf.py -m gpu -x -q -k "not full_depth" >> $L 2>&1
echo "tests rc=$?" >> $L
timeout 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench8.json 2>> $L
cat $L | cut -c1-200 | tail -20
