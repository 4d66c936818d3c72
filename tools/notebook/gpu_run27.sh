#!/bin/bash
# per-kernel device time + DRAM bytes of the decode step (eager launches; 4 steps x 195 kernels; skip the first 2 steps)
mkdir -p gpurun_out
N_NEW=5 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:'gemv_kernel|decode_' -s 390 -c 195 --csv --log-file gpurun_out/r27_decode_launches.csv \
  python tools/gpu_decode_prof.py > gpurun_out/r27.log 2>&1
tail -3 gpurun_out/r27.log; wc -l gpurun_out/r27_decode_launches.csv
