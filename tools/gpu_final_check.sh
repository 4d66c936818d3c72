# last check of the round on the final tree: whole GPU suite + smoke (no bench: profiles/r03d_* holds the evidence run)
mkdir -p gpurun_out
T=${TAG:-r03h}
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
python -c "
import json
d = json.load(open('gpurun_out/${T}_bench_n1.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['stages']['roofline_by_stage'])
"
