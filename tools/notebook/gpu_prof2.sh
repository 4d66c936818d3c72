#!/bin/bash
# ncu evidence for the bench command (round 1, final state):
#  (1) per-launch durations of one full step (our kernels only), (2) --set full of the decoder GEMMs of one layer,
#  (3) --set full of one decoder attention + one ViT attention launch
mkdir -p gpurun_out
KR='regex:gemm_bf16_kernel|attn2_fwd_kernel|attn_fwd_kernel|norm_kernel|rope_kernel|rope_table_kernel|merge_|im2col_patch14|pixel_shuffle'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -s 455 -c 455 --csv \
    --log-file gpurun_out/launches_r01.csv python bench.py --quick --steps 1 $BENCH_EXTRA > gpurun_out/prof2.log 2>&1
echo "launches rc=$?" >> gpurun_out/prof2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 355 -c 4 \
    -o gpurun_out/prof_gemm_dec_r01 -f python bench.py --quick --steps 1 $BENCH_EXTRA >> gpurun_out/prof2.log 2>&1
echo "gemm rc=$?" >> gpurun_out/prof2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn2_fwd_kernel -s 86 -c 1 \
    -o gpurun_out/prof_attn_dec_r01 -f python bench.py --quick --steps 1 $BENCH_EXTRA >> gpurun_out/prof2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn2_fwd_kernel -s 60 -c 1 \
    -o gpurun_out/prof_attn_vit_r01 -f python bench.py --quick --steps 1 $BENCH_EXTRA >> gpurun_out/prof2.log 2>&1
echo "attn rc=$?" >> gpurun_out/prof2.log
tail -4 gpurun_out/prof2.log
