#!/bin/bash
# Round 2, second GPU call: SFT suite after the fixes + ncu --set full captures (with source) of the kernels VERDICT r1 names:
# the short-K ViT GEMMs with a residual epilogue and the ViT attention.
mkdir -p gpurun_out
T=${TAG:-r02b}
run() { local name=$1; shift; echo "== $name"; timeout "${TMO:-600}" "$@" > gpurun_out/${T}_${name}.log 2>&1; echo "   exit $?"; tail -3 gpurun_out/${T}_${name}.log; }
run train_tests python -m pytest tests/test_zz_gpu_train.py -q -rA --tb=short
NCU="ncu --set full --clock-control none --import-source on"
for c in v_out v_fc2 v_patch v_qkv; do
  TMO=300 run ncu_gemm_$c $NCU -k regex:gemm_bf16 -s 1 -c 1 -o gpurun_out/${T}_gemm_$c -f python tools/gpu_check_gemm.py --pair --case $c
done
TMO=300 run ncu_attn_vit $NCU -k regex:attn3 -s 1 -c 1 -o gpurun_out/${T}_attn_vit -f python tools/gpu_check_attn.py --case vit_28 --variant 80 --time
ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
