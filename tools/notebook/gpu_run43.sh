#!/bin/bash
mkdir -p gpurun_out
LEOPARD_PREFILL_PDL=1 timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/gpu_ab_prefill.py pdl=0 pdl=1 2>&1 | tail -2
