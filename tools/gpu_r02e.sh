#!/bin/bash
# Round 2, fifth GPU call: parity mode (fp32-equivalent forward) against the oracle, the whole GPU suite after the fused-norm
# invariance fix, N = 1 bench, then (separate call) N = 2.
mkdir -p gpurun_out
T=${TAG:-r02e}
run() { local name=$1; shift; echo "== $name"; timeout "${TMO:-600}" "$@" > gpurun_out/${T}_${name}.log 2>&1; echo "   exit $?"; tail -c 900 gpurun_out/${T}_${name}.log; echo; }
run parity_small python -m pytest tests/test_gpu_parity_mode.py -q -x -s -k "glue or split_gemm or reduced"
TMO=2400 run parity_full python -m pytest tests/test_gpu_parity_mode.py -q -s -k "full_depth"
TMO=1800 run suite python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_mode.py
ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
