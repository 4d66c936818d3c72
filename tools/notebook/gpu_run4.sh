#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run4.log
: > $L
timeout 900 python -m pytest tests/test_gpu_hf.py -m gpu -q -s >> $L 2>&1
echo "hf rc=$?" >> $L
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "smoke rc=$?" >> $L
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2>> $L
echo "bench rc=$?" >> $L
cat gpurun_out/bench1.json >> $L
tail -60 $L
