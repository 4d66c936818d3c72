#!/bin/bash
# Every GPU call of round 2 as it was issued (`gpurun -- 'bash tools/gpu_runs_r02.sh <run>'`), one shell function per call; the files
# under profiles/ carry the run's tag as prefix (r02a_* ... r02l_*).  The round's final evidence run is tools/gpu_evidence_r02.sh, the
# N-GPU runs tools/gpu_multi_r02.sh.  Kept as the record of how each number in DESIGN.md / profiles/ was produced.
mkdir -p gpurun_out
run() { local name=$1; shift; echo "== $name"; timeout "${TMO:-600}" "$@" > gpurun_out/${T}_${name}.log 2>&1; echo "   exit $?"; tail -c ${TAIL:-1200} gpurun_out/${T}_${name}.log; echo; }

round2_first() {
  # First GPU call of the next round: everything that was written after round 1's GPU budget ran out, in the order that
  # localises a failure fastest.  Usage:  gpurun --timeout 2400 -- 'bash tools/gpu_round2_first.sh'
  # Each stage runs in its own process (a trapped kernel poisons only its stage) and writes one log under gpurun_out/.
  T=${TAG:-r02a}
  ZZ=tests/test_zz_gpu_train.py
  # 1. per-kernel parity of lp_train.cu (element-wise / reduction kernels; emulation-green already)
  run kernels   python -m pytest $ZZ --runxfail -q -x -k "transpose or rmsnorm or swiglu or colsum or masked_loss or softmax_backward or grad_norm or vision_tower_backward_kernels"
  # 2. whole step, materialised attention backward, frozen and trainable tower, against autograd
  run step_mat  python -m pytest $ZZ --runxfail -q -x -k "whole_sft_step_against_autograd or trainable_vision_tower"
  # 3. flash backward kernels (tcgen05) — per case, so that one wedged configuration does not hide the others
  for k in "1-128-4-1-True" "2-200-8-2-True" "1-384-4-4-False" "1-1024-8-2-True"; do
    run flash_$k python -m pytest $ZZ --runxfail -q -k "closed_form and $k"
  done
  run flash_vit python -m pytest $ZZ --runxfail -q -k "vision_tower_layout"
  run step_flash python -m pytest $ZZ --runxfail -q -k "flash_backward_matches_materialised"
  # 4. first timings: 4 decoder layers, then the full model, both attention-backward routes
  TMO=900 run speed_l4_mat   python tools/gpu_train_speed.py --layers 4 --vit-layers 27 --steps 2
  LEOPARD_ATTN_BWD=flash TMO=900 run speed_l4_flash python tools/gpu_train_speed.py --layers 4 --vit-layers 27 --steps 2
  LEOPARD_ATTN_BWD=flash TMO=1500 run speed_full_flash python tools/gpu_train_speed.py --steps 2
  # recompute on / off at 16 layers (the full model + kept activations needs ~172 GB: too close to the 180 GB for a first run)
  LEOPARD_ATTN_BWD=flash TMO=900 run speed_l16_flash_full python tools/gpu_train_speed.py --layers 16 --steps 2
  LEOPARD_ATTN_BWD=flash TMO=900 run speed_l16_flash_none python tools/gpu_train_speed.py --layers 16 --steps 2 --recompute none
  # 4b. the 4-CTA multicast GEMM (lp_gemm4.cu): correctness sweep + timing against the pair kernel, then the prefill A/B
  TMO=600 run gemm_quad_eq python -m pytest $ZZ --runxfail -q -k four_cta_multicast
  TMO=900 run gemm_quad python tools/gpu_check_gemm.py --quad --time
  TMO=900 run gemm_pair python tools/gpu_check_gemm.py --pair --time
  TMO=900 run ab_quad python tools/gpu_ab_prefill.py pair=1 pair=2
  # 5. the example loop (Megatron-shim surface) and the regression suite
  TMO=900 run example python examples/llava_sft_loop.py --layers 2 --iters 4
  TMO=1500 run suite python -m pytest tests -m gpu -q -x
  # 6. launch list of one 4-layer step (share of GEMM / transpose / attention-backward / optimizer)
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${T}_sft_launches.csv \
      python tools/gpu_train_speed.py --layers 4 --vit-layers 2 --steps 0 > gpurun_out/${T}_sft_prof.log 2>&1
  python tools/ncu_launch_agg.py gpurun_out/${T}_sft_launches.csv > gpurun_out/${T}_sft_launch_agg.txt 2>&1; head -30 gpurun_out/${T}_sft_launch_agg.txt
  ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
}

r02a() {
  # Round 2, first GPU call: what round 1 wrote after its GPU budget ran out (SFT step, flash backward, 4-CTA GEMM).
  # Usage: gpurun --timeout 1800 -- 'bash tools/gpu_r02a.sh'
  T=${TAG:-r02a}
  ZZ=tests/test_zz_gpu_train.py
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.log 2>&1
  run train_tests python -m pytest $ZZ --runxfail -q -rA --tb=long
  TMO=600 run speed_l4_mat   python tools/gpu_train_speed.py --layers 4 --vit-layers 27 --steps 2
  LEOPARD_ATTN_BWD=flash TMO=600 run speed_l4_flash python tools/gpu_train_speed.py --layers 4 --vit-layers 27 --steps 2
  LEOPARD_ATTN_BWD=flash TMO=900 run speed_full_flash python tools/gpu_train_speed.py --steps 2
  LEOPARD_ATTN_BWD=flash TMO=600 run speed_l16_flash_none python tools/gpu_train_speed.py --layers 16 --steps 2 --recompute none
  TMO=600 run gemm_quad python tools/gpu_check_gemm.py --quad --time
  TMO=600 run gemm_pair python tools/gpu_check_gemm.py --pair --time
  TMO=600 run ab_quad python tools/gpu_ab_prefill.py pair=1 pair=2
  ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
}

r02b() {
  # Round 2, second GPU call: SFT suite after the fixes + ncu --set full captures (with source) of the kernels VERDICT r1 names:
  # the short-K ViT GEMMs with a residual epilogue and the ViT attention.
  T=${TAG:-r02b}
  run train_tests python -m pytest tests/test_zz_gpu_train.py -q -rA --tb=short
  NCU="ncu --set full --clock-control none --import-source on"
  for c in v_out v_fc2 v_patch v_qkv; do
    TMO=300 run ncu_gemm_$c $NCU -k regex:gemm_bf16 -s 1 -c 1 -o gpurun_out/${T}_gemm_$c -f python tools/gpu_check_gemm.py --pair --case $c
  done
  TMO=300 run ncu_attn_vit $NCU -k regex:attn3 -s 1 -c 1 -o gpurun_out/${T}_attn_vit -f python tools/gpu_check_attn.py --case vit_28 --variant 80 --time
  ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
}

r02c() {
  # Round 2, third GPU call: new bench legs (library GPU competitor, strong scaling at N=1, cfg2 / cfg4 / cfg5 workloads, bounded CPU
  # arm), full-depth oracle fixtures for cfg2 / cfg3, ncu captures of the short-K ViT GEMMs.
  T=${TAG:-r02c}
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
}

r02d() {
  # Round 2, fourth GPU call: fused-norm / fused-RoPE GEMM epilogues — op parity, model parity, in-process A/B, whole GPU suite.
  T=${TAG:-r02d}
  run ops python -m pytest tests/test_gpu_ops.py -q -x -k "gemm_ex"
  run model python -m pytest tests/test_gpu_model.py -q -x -s -k "fused_norm or end_to_end or decoder or vit"
  run ab python tools/gpu_ab_prefill.py fused=0 fused=1
  TMO=1500 run suite python -m pytest tests -m gpu -q -x
  run bench_n1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
  ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
}

r02e() {
  # Round 2, fifth GPU call: parity mode (fp32-equivalent forward) against the oracle, the whole GPU suite after the fused-norm
  # invariance fix, N = 1 bench, then (separate call) N = 2.
  T=${TAG:-r02e}
  run parity_small python -m pytest tests/test_gpu_parity_mode.py -q -x -s -k "glue or split_gemm or reduced"
  TMO=2400 run parity_full python -m pytest tests/test_gpu_parity_mode.py -q -s -k "full_depth"
  TMO=1800 run suite python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_mode.py
  ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
}

r02f() {
  # Round 2, 2-GPU call: tensor-parallel fused kernels (all-gather -> GEMM, GEMM -> reduce-scatter over NVLink peer memory), the DP
  # scatter / gather check, and the N = 2 bench line (NCCL scatter / gather inside the timed step, strong-scaling queue).
  T=${TAG:-r02f}
  N=${NGPU:-2}
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  TMO=600 run tp_check $TR --master-port 29533 tests/gpu_tp_check.py
  TMO=600 run dp_check $TR --master-port 29534 tests/gpu_dp_check.py
  TMO=900 run bench_n$N $TR --master-port 29535 bench.py --gpus $N --steps 10 --warmup 3
  ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
}

r02h() {
  # Round 2, GEMM scheduling call: stream-K tail (accumulator hand-over between CTA pairs), residual tile by TMA, pipelined bias / column
  # sums, cached row statistics.  Parity first (bit-equality of every switch combination), then per-shape timing, then the whole step.
  T=${TAG:-r02h}
  TMO=900 run ops python -m pytest tests/test_gpu_ops.py -x -q -k "gemm"
  SH="--case d_o --case d_down --case d_qkv --case d_gateup --case v_out --case v_fc2 --case v_patch --case v_qkv --case v_fc1 --case p_1 --case dec_o --case dec_down"
  TAIL=4000 run gemm_new python tools/gpu_check_gemm.py --pair --time $SH
  TAIL=4000 run gemm_old python tools/gpu_check_gemm.py --pair --time --streamk 0 --tma-res 0 $SH
  TMO=900 run ab python tools/gpu_ab_prefill.py sk=0 sk=1 tmares=0 tmares=1
  TMO=900 TAIL=6000 run bench_n1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras
  TMO=900 run model python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -x -q
  ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
}

r02i() {
  # Round 2: stream-K with the data-parallel waves first (START / whole / END pieces at the end), every switch combination per shape,
  # whole-step A/B, ncu source-level capture of the LayerNorm-consuming fc1 GEMM inside a real step.
  T=${TAG:-r02i}
  TMO=900 run ops python -m pytest tests/test_gpu_ops.py -x -q -k "gemm"
  SH="--case d_o --case d_down --case d_qkv --case dec_o --case dec_down --case dec_qkv --case p_1 --case v_fc2 --case v_out"
  for sk in 0 1; do for tm in 0 1; do
    TAIL=3000 run gemm_sk${sk}_tma${tm} python tools/gpu_check_gemm.py --pair --time --streamk $sk --tma-res $tm $SH
  done; done
  TMO=900 run ab python tools/gpu_ab_prefill.py sk=0 sk=1
  NCU="ncu --set full --clock-control none --import-source on"
  TMO=400 run ncu_fc1 $NCU -k regex:gemm_bf16_kernel --launch-skip 122 --launch-count 1 -o gpurun_out/${T}_gemm_fc1_ln -f python bench.py --quick --steps 1 --warmup 1
  ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
}

r02j() {
  # Round 2: chained single-kernel projector, padded-batch HF merge / batched generate, full GPU suite after the GEMM scheduling work.
  T=${TAG:-r02j}
  TMO=600 run new python -m pytest tests/test_gpu_model.py tests/test_gpu_hf.py -x -q -k "projector or padded"
  TMO=600 run ab python tools/gpu_ab_prefill.py chain=0 chain=1
  TMO=1500 run suite python -m pytest tests -m gpu -x -q
  ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
}

r02l() {
  # Round 2: where the SFT step's time goes (ncu launch list of one cfg5 step, trainable tower) + the other workloads on the new GEMM code.
  T=${TAG:-r02l}
  TMO=900 run sft_speed python tools/gpu_train_speed.py --steps 3
  TMO=1200 run sft_launches ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_sft_launches.csv python tools/gpu_train_speed.py --steps 0
  TMO=600 run bench_cfg2 python bench.py --workload cfg2 --steps 10 --no-cpu-baseline --no-extras
  TMO=600 run bench_cfg4 python bench.py --workload cfg4 --steps 5
  TMO=900 run bench_cfg5 python bench.py --workload cfg5 --steps 3
  ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
}

r02o() {
  # Idefics2 GPU image processor + pure stream-K A/B (cfg3 and cfg2)
  T=${TAG:-r02o}
  run tests python -m pytest tests/test_gpu_frontend.py tests/test_gpu_idefics2.py tests/test_gpu_ops.py -x -q
  run ab_cfg3 python tools/gpu_ab_prefill.py sk=1 sk=2
  LEOPARD_AB_WORKLOAD=cfg2 run ab_cfg2 python tools/gpu_ab_prefill.py sk=1 sk=2
}

r02p() {
  # decode: L2 prefetch of the next projections' weights during the attention kernel
  T=${TAG:-r02p}
  TMO=900 run decode_prefetch python tools/gpu_decode_prefetch.py
  run tests python -m pytest tests/test_gpu_hf.py tests/test_gpu_decode.py -x -q
}

r02r() {
  # cfg5 bench line with explicit transposes vs MN-major operands (tower trainable), and frozen
  T=${TAG:-r02r}
  LEOPARD_SFT_OPERANDS=transpose run cfg5_transpose python bench.py --workload cfg5 --steps 3
  LEOPARD_SFT_OPERANDS=mn run cfg5_mn python bench.py --workload cfg5 --steps 3
  LEOPARD_SFT_OPERANDS=mn run cfg5_mn_frozen python bench.py --workload cfg5 --steps 3 --sft-freeze-vit
}

r02q() {
  # Round 2: transposed (MN-major) GEMM operands for the SFT step's dgrad / wgrad — parity, then the step with and without them.
  T=${TAG:-r02q}
  run ops python -m pytest tests/test_gpu_ops.py -x -q -k "transposed or stream_k"
  TMO=900 run train_tests python -m pytest tests/test_zz_gpu_train.py -x -q
  LEOPARD_SFT_OPERANDS=transpose TMO=600 run sft_transpose python tools/gpu_train_speed.py --steps 3
  LEOPARD_SFT_OPERANDS=mn TMO=600 run sft_mn python tools/gpu_train_speed.py --steps 3
  ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
}

r02v() {
  # attention v4 (lp_attn4.cu: one softmax warpgroup, two rows per thread as independent exp streams) vs v3: parity, bit-equality, timing,
  # then the whole cfg3 step with each
  T=${TAG:-r02v}
  TMO=240 run attn4_small python tools/gpu_check_attn.py --variant 128 --case dec_t128 --case dec_t128_c --case dec_t256 --case dec_t300_c --case vit_t128 --case vit_t676
  TMO=300 TAIL=4000 run attn4_time python tools/gpu_check_attn.py --variant 80 --variant 128 --variant 132 --case dec_2937_c --case dec_4796_c --case vit_28 --case vit_4900 --time
  TMO=300 run ab_attn python tools/gpu_ab_prefill.py attn=80 attn=128
}

r02x() {
  # attention v3 with the next S half loaded before the hand-over + alternating pv_done barriers: parity (v4 = the unchanged arithmetic as
  # cross-check), timing, phase trace, whole step
  T=${TAG:-r02x}
  TMO=240 run attn_small python tools/gpu_check_attn.py --variant 80 --variant 128 --case dec_t128 --case dec_t128_c --case dec_t256 --case dec_t300_c --case vit_t128 --case vit_t676
  TMO=300 TAIL=4000 run attn_time python tools/gpu_check_attn.py --variant 80 --variant 96 --variant 128 --case dec_2937_c --case dec_4796_c --case vit_28 --case vit_4900 --time
  timeout 60 tools/micro/attn_trace 4796 32 8 128 1 1 4 > gpurun_out/${T}_trace_dec_ping1.json 2>&1
  timeout 60 tools/micro/attn_trace 676 16 16 72 0 1 2 28 > gpurun_out/${T}_trace_vit_ping1.json 2>&1
  TMO=300 run ab_attn python tools/gpu_ab_prefill.py attn=80 attn=128
}

r02y() {
  # attention grids with the heads fastest (global heaviest-first dispatch of the causal query-tile pairs): parity, timing, whole step,
  # flash backward parity
  T=${TAG:-r02y}
  TMO=240 run attn_small python tools/gpu_check_attn.py --variant 80 --case dec_t128 --case dec_t128_c --case dec_t256 --case dec_t300_c --case vit_t128 --case vit_t676
  TMO=300 TAIL=4000 run attn_time python tools/gpu_check_attn.py --variant 80 --variant 128 --case dec_2937_c --case dec_4796_c --case vit_28 --case vit_4900 --time
  TMO=300 run ab_attn python tools/gpu_ab_prefill.py attn=80 attn=128
  TMO=600 run bwd python -m pytest tests/test_zz_gpu_train.py -x -q -k "flash or closed_form or vision_tower_layout"
}

r02z() {
  # GEMM raster: row-major tile order for ragged N with W L2-resident (ViT fc2): per-shape timing off / on, bit-equality, whole step
  T=${TAG:-r02z}
  TMO=300 run gemm_bal0 python tools/gpu_check_gemm.py --pair --time --balance 0 --case v_fc2 --case v_out --case v_fc1 --case v_qkv
  TMO=300 run gemm_bal1 python tools/gpu_check_gemm.py --pair --time --balance 1 --case v_fc2 --case v_out --case v_fc1 --case v_qkv
  TMO=300 run ab_bal python tools/gpu_ab_prefill.py bal=0 bal=1
  TMO=600 run ops python -m pytest tests/test_gpu_ops.py -x -q
}

r03a() {
  # persistent ViT attention (lp_attn5.cu, variant bit 8): parity + bit-equality with v3, timing, whole step
  T=${TAG:-r03a}
  TMO=240 run attn5_small python tools/gpu_check_attn.py --variant 336 --case vit_t676 --case vit_5x420 --case vit_2x1024 --case vit_17 --case vit_t128 --case dec_t300_c
  TMO=300 TAIL=4000 run attn5_time python tools/gpu_check_attn.py --variant 80 --variant 336 --variant 368 --case vit_28 --case vit_17 --case vit_4900 --time
  TMO=300 run ab_attn5 python tools/gpu_ab_prefill.py attn=80 attn=336
}

r03b() {
  # mid-session validation of the scheduling changes (attention dispatch order, persistent ViT attention, GEMM raster): whole GPU suite,
  # bench N=1 with the un-profiled stage timings, SFT step (flash backward grids)
  T=${TAG:-r03b}
  timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_pytest_gpu.log
  timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; tail -c 300 gpurun_out/${T}_bench_n1.err
  python -c "
import json
d = json.load(open('gpurun_out/${T}_bench_n1.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
print(d['stages']['roofline_by_stage'], d['stages'].get('stage_sum_ms'))
print(d['stages']['roofline_by_stage_kernel_events'])
print(d['roofline'])
print(d['decode'])
"
  timeout 900 python bench.py --workload cfg5 --steps 3 > gpurun_out/${T}_bench_cfg5.json 2> gpurun_out/${T}_bench_cfg5.err; tail -c 400 gpurun_out/${T}_bench_cfg5.err; head -c 600 gpurun_out/${T}_bench_cfg5.json
}

r03c() {
  # S-barrier probes (before and after the exponentials) in the attention kernels: timing, trace
  T=${TAG:-r03c}
  TMO=300 TAIL=3000 run attn_time python tools/gpu_check_attn.py --variant 336 --case dec_2937_c --case dec_4796_c --case vit_28 --case vit_4900 --case dec_t300_c --case vit_t676 --time
  timeout 60 tools/micro/attn_trace 4796 32 8 128 1 1 4 > gpurun_out/${T}_trace_dec.json 2>&1
  python tools/micro/attn_trace_report.py gpurun_out/${T}_trace_dec.json > gpurun_out/${T}_trace_report.txt; head -14 gpurun_out/${T}_trace_report.txt
}

r03f() {
  # power wall: do the polling epilogue warps cost clock?  nanosleep between probes of the accumulator-full barrier, whole-step A/B
  T=${TAG:-r03f}
  TMO=400 run ab_episleep python tools/gpu_ab_prefill.py episleep=0 episleep=200 episleep=1000 episleep=4000
}

r03e() {
  # N = 2 bench line on the last code (gpurun --gpus 2)
  mkdir -p gpurun_out
  T=r03e
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
  timeout 900 $TR --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_n2.log 2>&1
  grep '"metric"' gpurun_out/${T}_bench_n2.log > gpurun_out/${T}_bench_n2.json
  python -c "
  import json
  d=json.load(open('gpurun_out/r03e_bench_n2.json'))
  print(d['value'], d['ms_per_step'], d['e2e']['value'], d['dp'], d['strong']['ms_per_step'] if d.get('strong') else None)
  "
}

r03h() {
  # last check of the round on the final tree: whole GPU suite + smoke (no bench: profiles/r03d_* holds the evidence run)
  mkdir -p gpurun_out
  T=${TAG:-r03h}
  timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
  timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
  python -c "
  import json
  d = json.load(open('gpurun_out/${T}_bench_n1.json'))
  print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['stages']['roofline_by_stage'])
  "
}

r03j() {
  # compute-sanitizer memcheck over the kernels touched in the last session: persistent ViT attention (variant 336), v3 with the new grid
  # order / early S load, v4, and the GEMM with the balanced raster (ops tests run the fc2 shape)
  mkdir -p gpurun_out
  T=${TAG:-r03j}
  timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/gpu_check_attn.py --variant 336 --variant 128 \
      --case vit_t676 --case vit_5x420 --case dec_t300_c --case dec_t256 > gpurun_out/${T}_memcheck_attn.txt 2>&1
  tail -6 gpurun_out/${T}_memcheck_attn.txt
  timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_ops.py -q -x -k "raster_balance or scheduling_variants" > gpurun_out/${T}_memcheck_ops.txt 2>&1
  tail -6 gpurun_out/${T}_memcheck_ops.txt
}

r03k() {
  # N = 8 bench line on the last code (gpurun --gpus 8)
  mkdir -p gpurun_out
  T=r03k
  N=${NGPU:-8}
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  timeout 900 $TR --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_n$N.log 2>&1
  grep '"metric"' gpurun_out/${T}_bench_n$N.log > gpurun_out/${T}_bench_n$N.json
  python -c "
  import json
  d=json.load(open('gpurun_out/${T}_bench_n$N.json'))
  print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['dp'], d['strong']['ms_per_step'] if d.get('strong') else None, d['clocks'])
  "
}

"$@"
