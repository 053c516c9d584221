set -x
NG=${NG:-8}
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/bench_r2_n$NG.json 2> gpurun_out/bench_r2_n$NG.err ) 2>&1 | tail -4
tail -c 1200 gpurun_out/bench_r2_n$NG.err
python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/bench_r2_n$NG.json') if l.startswith('{')][-1])
for k in ('value','ms_per_step','c4_strong','c5_sharded'): print(k, json.dumps(d.get(k))[:2500])
print('e2e', d['e2e']['value'], d['e2e'].get('numa'))
"
