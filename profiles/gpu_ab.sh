timeout 800 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
bash profiles/gpu_quick.sh
