#!/bin/bash
# Round 2: chained single-kernel projector, padded-batch HF merge / batched generate, full GPU suite after the GEMM scheduling work.
mkdir -p gpurun_out
T=${TAG:-r02j}
run() { local name=$1; shift; echo "== $name"; timeout "${TMO:-600}" "$@" > gpurun_out/${T}_${name}.log 2>&1; echo "   exit $?"; tail -c ${TAIL:-1500} gpurun_out/${T}_${name}.log; echo; }
TMO=600 run new python -m pytest tests/test_gpu_model.py tests/test_gpu_hf.py -x -q -k "projector or padded"
TMO=600 run ab python tools/gpu_ab_prefill.py chain=0 chain=1
TMO=1500 run suite python -m pytest tests -m gpu -x -q
ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
