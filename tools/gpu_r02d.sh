#!/bin/bash
# Round 2, fourth GPU call: fused-norm / fused-RoPE GEMM epilogues — op parity, model parity, in-process A/B, whole GPU suite.
mkdir -p gpurun_out
T=${TAG:-r02d}
run() { local name=$1; shift; echo "== $name"; timeout "${TMO:-600}" "$@" > gpurun_out/${T}_${name}.log 2>&1; echo "   exit $?"; tail -c 700 gpurun_out/${T}_${name}.log; echo; }
run ops python -m pytest tests/test_gpu_ops.py -q -x -k "gemm_ex"
run model python -m pytest tests/test_gpu_model.py -q -x -s -k "fused_norm or end_to_end or decoder or vit"
run ab python tools/gpu_ab_prefill.py fused=0 fused=1
TMO=1500 run suite python -m pytest tests -m gpu -q -x
run bench_n1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
