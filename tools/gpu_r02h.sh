#!/bin/bash
# Round 2, GEMM scheduling call: stream-K tail (accumulator hand-over between CTA pairs), residual tile by TMA, pipelined bias / column
# sums, cached row statistics.  Parity first (bit-equality of every switch combination), then per-shape timing, then the whole step.
mkdir -p gpurun_out
T=${TAG:-r02h}
run() { local name=$1; shift; echo "== $name"; timeout "${TMO:-600}" "$@" > gpurun_out/${T}_${name}.log 2>&1; echo "   exit $?"; tail -c ${TAIL:-1200} gpurun_out/${T}_${name}.log; echo; }
TMO=900 run ops python -m pytest tests/test_gpu_ops.py -x -q -k "gemm"
SH="--case d_o --case d_down --case d_qkv --case d_gateup --case v_out --case v_fc2 --case v_patch --case v_qkv --case v_fc1 --case p_1 --case dec_o --case dec_down"
TAIL=4000 run gemm_new python tools/gpu_check_gemm.py --pair --time $SH
TAIL=4000 run gemm_old python tools/gpu_check_gemm.py --pair --time --streamk 0 --tma-res 0 $SH
TMO=900 run ab python tools/gpu_ab_prefill.py sk=0 sk=1 tmares=0 tmares=1
TMO=900 TAIL=6000 run bench_n1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras
TMO=900 run model python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -x -q
ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
