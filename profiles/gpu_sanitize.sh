timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/sanitize2.py > gpurun_out/sanitize2_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/sanitize2_racecheck.log
timeout 600 python -m pytest tests/test_pose_gpu.py tests/test_motion_gpu.py -x -q -m gpu 2>&1 | tail -2
