# round-end style validation: GPU tests, smoke, both bench arms
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-600
python bench.py 2>&1 | tail -1 > gpurun_out/bench_final.json; python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); r=d['roofline']
print('value %.4g e2e %.4g ms/step %.4f frac %.3f launches %d clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['frac'], d['gpu_launches'], d['clocks']))
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['mask_equals_gpu'])
for k,v in d['secondary'].items(): print(k, v)"
