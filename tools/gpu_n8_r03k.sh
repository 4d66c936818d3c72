mkdir -p gpurun_out
T=r03k
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_n$N.log 2>&1
grep '"metric"' gpurun_out/${T}_bench_n$N.log > gpurun_out/${T}_bench_n$N.json
python -c "
import json
d=json.load(open('gpurun_out/${T}_bench_n$N.json'))
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['dp'], d['strong']['ms_per_step'] if d.get('strong') else None, d['clocks'])
"
