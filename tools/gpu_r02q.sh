#!/bin/bash
# Round 2: transposed (MN-major) GEMM operands for the SFT step's dgrad / wgrad — parity, then the step with and without them.
mkdir -p gpurun_out
T=${TAG:-r02q}
run() { local name=$1; shift; echo "== $name"; timeout "${TMO:-600}" "$@" > gpurun_out/${T}_${name}.log 2>&1; echo "   exit $?"; tail -c ${TAIL:-1200} gpurun_out/${T}_${name}.log; echo; }
run ops python -m pytest tests/test_gpu_ops.py -x -q -k "transposed or stream_k"
TMO=900 run train_tests python -m pytest tests/test_zz_gpu_train.py -x -q
LEOPARD_SFT_OPERANDS=transpose TMO=600 run sft_transpose python tools/gpu_train_speed.py --steps 3
LEOPARD_SFT_OPERANDS=mn TMO=600 run sft_mn python tools/gpu_train_speed.py --steps 3
ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
