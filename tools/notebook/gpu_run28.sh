#!/bin/bash
# block-cooperative GEMV + fused attention/combine: tests, speed per mode, ncu launch list of the new default
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_hf.py -x -q -s 2>&1 | tail -30 > gpurun_out/r28_tests.log
tail -4 gpurun_out/r28_tests.log
timeout 600 python tools/gpu_decode_speed.py > gpurun_out/r28_decode_speed.log 2>&1
tail -6 gpurun_out/r28_decode_speed.log
N_NEW=5 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:'gemv|decode_' -s 326 -c 163 --csv --log-file gpurun_out/r28_decode_launches.csv \
  python tools/gpu_decode_prof.py > gpurun_out/r28.log 2>&1
tail -2 gpurun_out/r28.log
