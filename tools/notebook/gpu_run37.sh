#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_hf.py -x -q 2>&1 | tail -4
timeout 300 python tools/gpu_decode_stamps.py 2>&1 | tail -8
timeout 600 python tools/gpu_decode_speed.py > gpurun_out/r37_decode_speed.log 2>&1
tail -8 gpurun_out/r37_decode_speed.log | cut -c1-120
