# compute-sanitizer memcheck over the kernels touched in the last session: persistent ViT attention (variant 336), v3 with the new grid
# order / early S load, v4, and the GEMM with the balanced raster (ops tests run the fc2 shape)
mkdir -p gpurun_out
T=${TAG:-r03j}
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/gpu_check_attn.py --variant 336 --variant 128 \
    --case vit_t676 --case vit_5x420 --case dec_t300_c --case dec_t256 > gpurun_out/${T}_memcheck_attn.txt 2>&1
tail -6 gpurun_out/${T}_memcheck_attn.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_ops.py -q -x -k "raster_balance or scheduling_variants" > gpurun_out/${T}_memcheck_ops.txt 2>&1
tail -6 gpurun_out/${T}_memcheck_ops.txt
