#!/bin/bash
# Round 2: stream-K with the data-parallel waves first (START / whole / END pieces at the end), every switch combination per shape,
# whole-step A/B, ncu source-level capture of the LayerNorm-consuming fc1 GEMM inside a real step.
mkdir -p gpurun_out
T=${TAG:-r02i}
run() { local name=$1; shift; echo "== $name"; timeout "${TMO:-600}" "$@" > gpurun_out/${T}_${name}.log 2>&1; echo "   exit $?"; tail -c ${TAIL:-1200} gpurun_out/${T}_${name}.log; echo; }
TMO=900 run ops python -m pytest tests/test_gpu_ops.py -x -q -k "gemm"
SH="--case d_o --case d_down --case d_qkv --case dec_o --case dec_down --case dec_qkv --case p_1 --case v_fc2 --case v_out"
for sk in 0 1; do for tm in 0 1; do
  TAIL=3000 run gemm_sk${sk}_tma${tm} python tools/gpu_check_gemm.py --pair --time --streamk $sk --tma-res $tm $SH
done; done
TMO=900 run ab python tools/gpu_ab_prefill.py sk=0 sk=1
NCU="ncu --set full --clock-control none --import-source on"
TMO=400 run ncu_fc1 $NCU -k regex:gemm_bf16_kernel --launch-skip 122 --launch-count 1 -o gpurun_out/${T}_gemm_fc1_ln -f python bench.py --quick --steps 1 --warmup 1
ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
