mkdir -p gpurun_out
timeout 60 tools/micro/attn_trace 4796 32 8 128 1 1 4 > gpurun_out/r02w_trace_dec_ping1.json 2>&1
timeout 60 tools/micro/attn_trace 4796 32 8 128 1 0 4 > gpurun_out/r02w_trace_dec_ping0.json 2>&1
timeout 60 tools/micro/attn_trace 676 16 16 72 0 1 2 28 > gpurun_out/r02w_trace_vit_ping1.json 2>&1
timeout 60 tools/micro/attn_trace 4900 16 16 72 0 1 4 8 > gpurun_out/r02w_trace_vit4900_ping1.json 2>&1
ls -la gpurun_out/r02w_*; head -c 300 gpurun_out/r02w_trace_dec_ping1.json
