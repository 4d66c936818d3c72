#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_hf.py -x -q -s 2>&1 | tail -12 > gpurun_out/r36_tests.log
tail -6 gpurun_out/r36_tests.log
timeout 300 python tools/gpu_decode_stamps.py 2>&1 | tail -8
timeout 600 python tools/gpu_decode_speed.py > gpurun_out/r36_decode_speed.log 2>&1
tail -8 gpurun_out/r36_decode_speed.log | cut -c1-120
