#!/bin/bash
# Round 2, third GPU call: new bench legs (library GPU competitor, strong scaling at N=1, cfg2 / cfg4 / cfg5 workloads, bounded CPU
# arm), full-depth oracle fixtures for cfg2 / cfg3, ncu captures of the short-K ViT GEMMs.
mkdir -p gpurun_out
T=${TAG:-r02c}
run() { local name=$1; shift; echo "== $name"; timeout "${TMO:-600}" "$@" > gpurun_out/${T}_${name}.log 2>&1; echo "   exit $?"; tail -c 600 gpurun_out/${T}_${name}.log; echo; }
run bench_n1 python bench.py --steps 10 --warmup 3
run bench_ref python bench.py --impl reference --steps 3 --warmup 1
run bench_cfg2 python bench.py --workload cfg2 --steps 10 --no-cpu-baseline
run bench_cfg4 python bench.py --workload cfg4 --steps 5
run bench_cfg5 python bench.py --workload cfg5 --steps 3
run bench_cfg5_frozen python bench.py --workload cfg5 --steps 3 --sft-freeze-vit
TMO=1500 run fullsize python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_model.py -q -s -k "full_depth" 
NCU="ncu --set full --clock-control none --import-source on"
for c in v_out v_fc2 v_patch d_o; do
  TMO=300 run ncu_gemm_$c $NCU -k regex:gemm_bf16 -s 2 -c 1 -o gpurun_out/${T}_gemm_$c -f python tools/gpu_check_gemm.py --pair --time --case $c
done
ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
