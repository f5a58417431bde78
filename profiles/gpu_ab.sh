timeout 800 python -m pytest tests/test_pose_gpu.py tests/test_motion_gpu.py -x -q -m gpu 2>&1 | tail -3
bash profiles/gpu_quick.sh
