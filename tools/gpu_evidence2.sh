#!/bin/bash
# Final-state evidence (after prefill PDL): GPU suite, smoke, example, both bench arms, prefill launch list
mkdir -p gpurun_out
T=${TAG:-ev}
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "parity|passed|failed|error" > gpurun_out/${T}_pytest_gpu.log
tail -2 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 600 python examples/llava_eval_loop.py --synthetic --layers 2 --max-new-tokens 6 > gpurun_out/${T}_example.log 2>&1; tail -2 gpurun_out/${T}_example.log
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
tail -c 1200 gpurun_out/${T}_bench_n1.json
KR='regex:gemm_bf16_kernel|attn3_fwd_kernel|attn2_fwd_kernel|attn_fwd_kernel|norm_kernel|rmsnorm_|rope_kernel|rope_table_kernel|merge_|im2col_patch14|pixel_shuffle'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -s 455 -c 455 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --quick --steps 1 > gpurun_out/${T}_prof.log 2>&1
ls -la gpurun_out/${T}_* | awk '{print $5, $9}'
