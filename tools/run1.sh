python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_rev.json 2> gpurun_out/bench_rev.err; tail -c 300 gpurun_out/bench_rev.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_rev.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])
print(json.dumps(d['roofline'])[:800])
"
python tools/timeline.py 30 2>&1 | cut -c1-250 > gpurun_out/timeline_rev.txt; head -3 gpurun_out/timeline_rev.txt
for n in 24 26 28; do python tools/quick.py --m 3 --n $n --batch 4096 --iters 10 --sweep whole_order=104,232 2>&1 | tail -2; done
python tools/quick.py --m 2 --n 30 --batch 16384 --iters 10 --sweep whole_order=104,232 2>&1 | tail -2
python tools/quick.py --m 3 --n 20 --batch 8192 --iters 10 --sweep whole_order=104,232 2>&1 | tail -2
