#!/bin/bash
# compute-sanitizer memcheck over the new byte/stream kernels (image front-end, decode GEMV/attention)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_decode.py -x -q -k "not full_sizes" > gpurun_out/r38_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r38_memcheck.log | head -20
