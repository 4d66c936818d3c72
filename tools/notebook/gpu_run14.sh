#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/run14.log
: > $L
nvidia-smi -L >> $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/gpu_dp_check.py >> $L 2>&1
echo "dp rc=$?" >> $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r01_n2.json 2>> $L
echo "bench2 rc=$?" >> $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_r01_n2_ref.json 2>> $L
echo "ref2 rc=$?" >> $L
tail -15 $L; cat gpurun_out/bench_r01_n2.json | head -c 400
