#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_idefics2_cfg4.py > gpurun_out/idefics2_cfg4.json 2> gpurun_out/run21.err
echo "rc=$?"; cat gpurun_out/idefics2_cfg4.json; tail -5 gpurun_out/run21.err
