set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py > gpurun_out/bench_ours_final.json 2> gpurun_out/bench_ours_final.err ) 2>&1 | tail -4
tail -c 1500 gpurun_out/bench_ours_final.err
python -c "
import json
d=json.load(open('gpurun_out/bench_ours_final.json'))
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'traffic',d['roofline'].get('traffic'))
print('e2e',d['e2e']['value'],d['e2e'].get('pcie_frac'),d['e2e']['pageable'])
print('cpu',d['cpu_baseline'])
print('d6',d['d6_single']['ms_per_roundtrip'],d['d6_single']['plan_device_MB'],d['d6_single']['plan_build_s'],d['d6_single']['eval_10_points_ms'])
print('c5',d['c5_dx_eval']['eval_ms'],d['c5_dx_eval']['dx_us'], d['c5_dx_eval']['roofline']['frac'])
print('clocks',d['clocks'],'launches',d['gpu_launches'])
"
( time python bench.py --impl reference > gpurun_out/bench_reference_final.json 2> gpurun_out/bench_reference_final.err ) 2>&1 | tail -4
cat gpurun_out/bench_reference_final.json | cut -c1-400
