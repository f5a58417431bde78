timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --workload c5 --steps 50 2>&1 | tail -1 > gpurun_out/bench_c5_n1.json
python -c "
import json; d=json.load(open('gpurun_out/bench_c5_n1.json')); r=d['roofline']
print('C5 N=1 value %.4g e2e %.4g ms/step %.4f | classify %.3f warp %.3f group %.3f | frac %.3f | B/pose %.0f | cpu %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['classify_kernel_ms'], r['kernel_ms'], r['group_kernel_ms'], r['frac'], r['algorithmic_bytes_per_pose'], d['cpu_baseline']))"
timeout 300 compute-sanitizer --tool racecheck --print-limit 3 python profiles/sanitize.py 2>&1 | grep -E "RACECHECK SUMMARY|hazard" | head -5
