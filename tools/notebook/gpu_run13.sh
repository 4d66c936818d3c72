#!/bin/bash
# round-1 evidence run: full GPU test-suite, smoke, bench (N=1) with CPU baseline, reference arm, ncu profiles
mkdir -p gpurun_out
L=gpurun_out/run13.log
: > $L
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_r01.log 2>&1
echo "pytest rc=$?" >> $L
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "smoke rc=$?" >> $L
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_n1.json 2>> $L
echo "bench rc=$?" >> $L
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r01_reference.json 2>> $L
echo "ref rc=$?" >> $L
bash tools/gpu_prof2.sh >> $L 2>&1
tail -12 $L
