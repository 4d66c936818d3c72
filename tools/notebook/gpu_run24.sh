#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run24.log
: > $L
timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "norm" >> $L 2>&1
echo "norm tests rc=$?" >> $L
timeout 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench11.json 2>> $L
python -c "
import json; d=json.load(open('gpurun_out/bench11.json')); print(d['value'], d['ms_per_step'], d['stages']['kernel_ms'])" >> $L
tail -6 $L
