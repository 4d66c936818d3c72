#!/bin/bash
# Round 2, N-GPU call (NGPU = 2 / 4 / 8): cfg3 bench line (NCCL scatter / gather inside the timed step + strong-scaling queue) and the cfg5
# SFT step with the bucketed NCCL gradient all-reduce (exposed communication time = step with / without the reduce).
mkdir -p gpurun_out
T=${TAG:-r02n}
N=${NGPU:-2}
run() { local name=$1; shift; echo "== $name"; timeout "${TMO:-900}" "$@" > gpurun_out/${T}_${name}.log 2>&1; echo "   exit $?"; tail -c 1800 gpurun_out/${T}_${name}.log; echo; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run bench_n$N $TR --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3
run cfg5_n$N $TR --master-port 29542 bench.py --workload cfg5 --gpus $N --steps 3
ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
