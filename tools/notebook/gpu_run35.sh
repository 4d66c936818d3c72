#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py -x -q -s 2>&1 | tail -25 > gpurun_out/r35_tests.log
tail -12 gpurun_out/r35_tests.log
timeout 600 python tools/gpu_decode_speed.py > gpurun_out/r35_decode_speed.log 2>&1
tail -8 gpurun_out/r35_decode_speed.log | cut -c1-200
