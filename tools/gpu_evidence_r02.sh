#!/bin/bash
# Round-2 evidence on one B200: GPU test suite, smoke, example, both bench arms, ncu launch list of one bench step, ncu --set full of
# the dominant GEMM (DRAM traffic for bench.py's roofline.traffic), the residual-ring GEMM, the chained projector and the v3 attention.
mkdir -p gpurun_out
T=${TAG:-r02m}
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "parity|passed|failed|error" > gpurun_out/${T}_pytest_gpu.log
tail -2 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 600 python examples/llava_eval_loop.py --synthetic --layers 2 --max-new-tokens 6 > gpurun_out/${T}_example.log 2>&1; tail -2 gpurun_out/${T}_example.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
tail -c 600 gpurun_out/${T}_bench_reference.json
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
tail -c 1500 gpurun_out/${T}_bench_n1.json
N=$(timeout 600 python bench.py --quick --steps 1 --warmup 1 2>/dev/null | python -c "import json,sys; print(json.loads(sys.stdin.readline())['gpu_launches'])")
echo "launches per step: $N"
KR='regex:gemm_|attn|norm_kernel|rmsnorm|rope|merge_|im2col|pixel_shuffle|row_stats'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -s $N -c $N --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --quick --steps 1 --warmup 1 > gpurun_out/${T}_prof.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:gemm_bf16_kernel -s 2 -c 1 -f -o gpurun_out/${T}_gemm_gateup python tools/gpu_check_gemm.py --pair --time --case d_gateup >> gpurun_out/${T}_prof.log 2>&1
timeout 600 $NCU -k regex:gemm_bf16_res_kernel -s 2 -c 1 -f -o gpurun_out/${T}_gemm_d_down python tools/gpu_check_gemm.py --pair --time --case d_down >> gpurun_out/${T}_prof.log 2>&1
timeout 600 $NCU -k regex:gemm_bf16_res_kernel -s 2 -c 1 -f -o gpurun_out/${T}_gemm_v_out python tools/gpu_check_gemm.py --pair --time --case v_out >> gpurun_out/${T}_prof.log 2>&1
timeout 600 $NCU -k regex:gemm_chain_kernel -s 1 -c 1 -f -o gpurun_out/${T}_gemm_chain python tools/gpu_ab_prefill.py chain=1 >> gpurun_out/${T}_prof.log 2>&1
timeout 600 $NCU -k regex:attn3_fwd -s 2 -c 1 -f -o gpurun_out/${T}_attn3_dec python tools/gpu_check_attn.py --case dec_4796_c --variant 80 --time >> gpurun_out/${T}_prof.log 2>&1
timeout 600 $NCU -k regex:attn3_fwd -s 2 -c 1 -f -o gpurun_out/${T}_attn3_vit python tools/gpu_check_attn.py --case vit_28 --variant 80 --time >> gpurun_out/${T}_prof.log 2>&1
timeout 600 $NCU -k regex:attn5_fwd -s 2 -c 1 -f -o gpurun_out/${T}_attn5_vit python tools/gpu_check_attn.py --case vit_28 --variant 336 --time >> gpurun_out/${T}_prof.log 2>&1
timeout 600 $NCU -k regex:gemm_bf16_res_kernel -s 2 -c 1 -f -o gpurun_out/${T}_gemm_v_fc2 python tools/gpu_check_gemm.py --pair --time --case v_fc2 >> gpurun_out/${T}_prof.log 2>&1
timeout 900 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/${T}_bench_cfg4.json 2> gpurun_out/${T}_bench_cfg4.err
timeout 900 python bench.py --workload cfg5 --steps 3 > gpurun_out/${T}_bench_cfg5.json 2> gpurun_out/${T}_bench_cfg5.err
for r in gemm_gateup gemm_d_down gemm_v_out gemm_v_fc2 gemm_chain attn3_dec attn3_vit attn5_vit; do
  python tools/ncu_summary.py gpurun_out/${T}_$r.ncu-rep > gpurun_out/${T}_$r.txt 2>/dev/null
done
rm -f gpurun_out/${T}_gemm_gateup.ncu-rep gpurun_out/${T}_gemm_d_down.ncu-rep gpurun_out/${T}_gemm_chain.ncu-rep gpurun_out/${T}_gemm_v_fc2.ncu-rep gpurun_out/${T}_attn3_vit.ncu-rep   # summaries travel; keep the reports with source pages
ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
