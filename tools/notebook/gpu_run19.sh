#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run19.log
: > $L
timeout 150 python -m pytest tests/test_gpu_hf.py -m gpu -x -q -k "generate or decode" >> $L 2>&1
echo "decode tests rc=$?" >> $L
timeout 120 python tools/gpu_decode_speed.py >> $L 2>&1
echo "decode speed rc=$?" >> $L
timeout 150 python tools/gpu_check_attn.py --variant 6 --case dec_t128_c --case dec_t300_c --case vit_t676 --case dec_t256 --case vit_t128 >> $L 2>&1
echo "v6 rc=$?" >> $L
timeout 200 python tools/gpu_check_attn.py --variant 6 --variant 2 --time --case dec_2937_c --case dec_4796_c --case vit_28 >> $L 2>&1
echo "attn rc=$?" >> $L
timeout 600 python -m pytest tests -m gpu -x -q -k "not full_depth" >> $L 2>&1
echo "tests rc=$?" >> $L
timeout 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench9.json 2>> $L
cat $L | cut -c1-200 | tail -22
