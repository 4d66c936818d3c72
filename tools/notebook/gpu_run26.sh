#!/bin/bash
# round-1 session 2, call 1: new decode + front-end tests, decode speed per implementation
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_frontend.py -x -q -s 2>&1 | tail -60 > gpurun_out/r26_tests.log
timeout 600 python tools/gpu_decode_speed.py > gpurun_out/r26_decode_speed.log 2>&1
tail -5 gpurun_out/r26_tests.log; cat gpurun_out/r26_decode_speed.log | tail -8
