#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_ops.py tests/test_gpu_model.py -x -q 2>&1 | tail -12
timeout 600 python tools/gpu_ab_prefill.py tail=0 tail=1 2>&1 | tail -2
