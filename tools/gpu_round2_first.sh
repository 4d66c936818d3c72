#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget ran out, in the order that
# localises a failure fastest.  Usage:  gpurun --timeout 2400 -- 'bash tools/gpu_round2_first.sh'
# Each stage runs in its own process (a trapped kernel poisons only its stage) and writes one log under gpurun_out/.
mkdir -p gpurun_out
T=${TAG:-r02a}
run() { local name=$1; shift; echo "== $name"; timeout "${TMO:-600}" "$@" > gpurun_out/${T}_${name}.log 2>&1; echo "   exit $?"; tail -3 gpurun_out/${T}_${name}.log; }
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
