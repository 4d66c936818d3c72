#!/bin/bash
# full GPU suite with the new defaults (attention v3 + turns, decode mode 5) and a bench run
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r34_tests.log
tail -5 gpurun_out/r34_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r34_bench.json 2> gpurun_out/r34_bench.err
tail -c 1500 gpurun_out/r34_bench.json
