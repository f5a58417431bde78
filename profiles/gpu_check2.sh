timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python bench.py --steps 50 2>&1 | tail -1 > gpurun_out/bench_latest.json; python -c "
import json; d=json.load(open('gpurun_out/bench_latest.json')); r=d['roofline']
print('value %.4g e2e %.4g ms/step %.4f frac %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['frac']))
print(json.dumps(d['secondary']['fused_sample_check_compact']))"
