python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_evaltree.json 2> gpurun_out/bench_evaltree.err; tail -c 300 gpurun_out/bench_evaltree.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_evaltree.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])
print(json.dumps(d.get('c5_dx_eval'))[:600])
print(json.dumps(d.get('d6_single'))[:600])
"
python tools/quick_eval.py 4 20 1e7 chebyshev
