set -x
( time python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err ) 2>&1 | tail -4
tail -c 3000 gpurun_out/bench_r2a.err
python -c "
import json
d=json.load(open('gpurun_out/bench_r2a.json'))
for k in ('value','ms_per_step','e2e','cpu_baseline','c4_strong','c5_sharded'): print(k, json.dumps(d.get(k))[:900])
print('d6', d['d6_single']['ms_per_roundtrip'], d['d6_single']['plan_device_MB'], 'c5', d['c5_dx_eval']['eval_ms'], d['c5_dx_eval']['dx_us'])
"
( time python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_ref_r2a.json 2> gpurun_out/bench_ref_r2a.err ) 2>&1 | tail -4
cat gpurun_out/bench_ref_r2a.json; tail -c 1000 gpurun_out/bench_ref_r2a.err
