# N = 1, 2, 4 weak-scaling runs of bench.py (one box, gpurun --gpus 4)
python bench.py --steps 100 2>&1 | tail -1 > gpurun_out/bench_n1.json
for N in 2 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 100 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n$N.json
done
for N in 1 2 4; do python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print('N=$N value %.4g ms/step %.4f e2e %.4g exchange_ok %s parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d.get('exchange_ok'), d['cpu_baseline']['mask_equals_gpu']))" || tail -5 gpurun_out/bench_n$N.json; done
