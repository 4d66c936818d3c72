#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -x -q 2>&1 | tail -3
timeout 600 python tools/gpu_ab_prefill.py norm=0 norm=1 2>&1 | tail -3
