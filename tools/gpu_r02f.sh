#!/bin/bash
# Round 2, 2-GPU call: tensor-parallel fused kernels (all-gather -> GEMM, GEMM -> reduce-scatter over NVLink peer memory), the DP
# scatter / gather check, and the N = 2 bench line (NCCL scatter / gather inside the timed step, strong-scaling queue).
mkdir -p gpurun_out
T=${TAG:-r02f}
N=${NGPU:-2}
run() { local name=$1; shift; echo "== $name"; timeout "${TMO:-600}" "$@" > gpurun_out/${T}_${name}.log 2>&1; echo "   exit $?"; tail -c 1500 gpurun_out/${T}_${name}.log; echo; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
TMO=600 run tp_check $TR --master-port 29533 tests/gpu_tp_check.py
TMO=600 run dp_check $TR --master-port 29534 tests/gpu_dp_check.py
TMO=900 run bench_n$N $TR --master-port 29535 bench.py --gpus $N --steps 10 --warmup 3
ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
