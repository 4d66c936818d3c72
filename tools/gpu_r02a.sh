#!/bin/bash
# Round 2, first GPU call: what round 1 wrote after its GPU budget ran out (SFT step, flash backward, 4-CTA GEMM).
# Usage: gpurun --timeout 1800 -- 'bash tools/gpu_r02a.sh'
mkdir -p gpurun_out
T=${TAG:-r02a}
run() { local name=$1; shift; echo "== $name"; timeout "${TMO:-600}" "$@" > gpurun_out/${T}_${name}.log 2>&1; echo "   exit $?"; tail -3 gpurun_out/${T}_${name}.log; }
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
