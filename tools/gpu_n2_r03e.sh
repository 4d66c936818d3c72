mkdir -p gpurun_out
T=r03e
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_n2.log 2>&1
grep '"metric"' gpurun_out/${T}_bench_n2.log > gpurun_out/${T}_bench_n2.json
python -c "
import json
d=json.load(open('gpurun_out/r03e_bench_n2.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['dp'], d['strong']['ms_per_step'] if d.get('strong') else None)
"
