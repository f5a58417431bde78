# N = 2 and 4 weak-scaling runs of bench.py (one box, gpurun --gpus 4)
for N in 2 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 100 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n$N.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print('N=$N value %.4g ms/step %.4f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done
