set -x
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err ) 2>&1 | tail -4
tail -c 1500 gpurun_out/bench_r2_n2.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_n2.json') if l.startswith('{')][-1])
for k in ('value','ms_per_step','c4_strong','c5_sharded','sharded_single'): print(k, json.dumps(d.get(k))[:1500])
print('e2e', d['e2e']['value'])
"
python -m pytest tests -m gpu -x -q -k "push or sharded" 2>&1 | tail -2
