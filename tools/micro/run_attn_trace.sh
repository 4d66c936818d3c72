mkdir -p gpurun_out
T=${TAG:-r03g}
timeout 60 tools/micro/attn_trace 4796 32 8 128 1 1 4 > gpurun_out/${T}_trace_dec.json 2>&1
python tools/micro/attn_trace_report.py gpurun_out/${T}_trace_dec.json > gpurun_out/${T}_trace_report.txt 2>&1; head -16 gpurun_out/${T}_trace_report.txt
