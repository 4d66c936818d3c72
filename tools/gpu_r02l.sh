#!/bin/bash
# Round 2: where the SFT step's time goes (ncu launch list of one cfg5 step, trainable tower) + the other workloads on the new GEMM code.
mkdir -p gpurun_out
T=${TAG:-r02l}
run() { local name=$1; shift; echo "== $name"; timeout "${TMO:-600}" "$@" > gpurun_out/${T}_${name}.log 2>&1; echo "   exit $?"; tail -c ${TAIL:-900} gpurun_out/${T}_${name}.log; echo; }
TMO=900 run sft_speed python tools/gpu_train_speed.py --steps 3
TMO=1200 run sft_launches ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_sft_launches.csv python tools/gpu_train_speed.py --steps 0
TMO=600 run bench_cfg2 python bench.py --workload cfg2 --steps 10 --no-cpu-baseline --no-extras
TMO=600 run bench_cfg4 python bench.py --workload cfg4 --steps 5
TMO=900 run bench_cfg5 python bench.py --workload cfg5 --steps 3
ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
