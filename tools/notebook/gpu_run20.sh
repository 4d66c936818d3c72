#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run20.log
: > $L
C="--case ragged --case ragged128 --case r224 --case r192 --case r160 --case sw192 --case sw128 --case fp32out --case vit_patch --case lm_head"
timeout 150 python tools/gpu_check_gemm.py --pair $C >> $L 2>&1
echo "pair small rc=$?" >> $L
timeout 150 python tools/gpu_check_gemm.py --single $C >> $L 2>&1
echo "single small rc=$?" >> $L
timeout 200 python tools/gpu_check_gemm.py --time --case v_out --case v_fc2 --case v_fc1 --case v_qkv --case d_o --case d_down >> $L 2>&1
echo "timing rc=$?" >> $L
timeout 600 python -m pytest tests -m gpu -x -q -k "not full_depth" >> $L 2>&1
echo "tests rc=$?" >> $L
timeout 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench10.json 2>> $L
python - <<'PY' >> $L
import json
for l in open('gpurun_out/run20.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print(d['case'], d['ok'], round(d.get('tflops',0)), round(d.get('cublas_tflops',0)), d.get('first_bad',''))
PY
tail -45 $L | cut -c1-160
