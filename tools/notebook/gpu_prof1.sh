#!/bin/bash
# ncu evidence for the bench command: (1) per-launch durations of one step, (2) full-set capture of the top kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 460 -c 460 --csv \
    --log-file gpurun_out/launches_r01.csv python bench.py --quick --steps 1 > gpurun_out/prof_bench.log 2>&1
echo "launches rc=$?" >> gpurun_out/prof_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 300 -c 4 \
    -o gpurun_out/prof_gemm_r01 -f python bench.py --quick --steps 1 >> gpurun_out/prof_bench.log 2>&1
echo "gemm rc=$?" >> gpurun_out/prof_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 40 -c 2 \
    -o gpurun_out/prof_attn_r01 -f python bench.py --quick --steps 1 >> gpurun_out/prof_bench.log 2>&1
echo "attn rc=$?" >> gpurun_out/prof_bench.log
timeout 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench2.json 2>> gpurun_out/prof_bench.log
tail -5 gpurun_out/prof_bench.log; ls -la gpurun_out/
