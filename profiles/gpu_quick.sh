python bench.py --steps 50 2>&1 | tail -1 > gpurun_out/bench_latest.json; python -c "
import json; d=json.load(open('gpurun_out/bench_latest.json')); r=d['roofline']
print('value %.4g e2e %.4g ms/step %.4f | classify %.3f warp %.3f group %.3f | frac %.3f | queued %d deferred %d parity %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['classify_kernel_ms'], r['kernel_ms'], r['group_kernel_ms'], r['frac'], r['queued_boxes'], r['deferred_boxes'], d['cpu_baseline']['mask_equals_gpu']))"
