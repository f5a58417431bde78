python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py 2>&1 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 3 2>&1 | tail -3
