python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_pruned.json 2> gpurun_out/bench_pruned.err; tail -c 400 gpurun_out/bench_pruned.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_pruned.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d.get('so_MB'), d.get('plan_device_MB'), d.get('plan_build_s'))
"
