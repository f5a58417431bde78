timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
bash profiles/gpu_quick.sh
