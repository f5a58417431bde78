#!/bin/bash
# r02 final code (v6): launch list of the default bench command, full captures of the pose-path kernels on the bench
# workload (group kernel, both one-warp-per-box launches, classify), a light capture of every motion-cost trunk kernel,
# compute-sanitizer over a small batch
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 140 --csv --log-file gpurun_out/launches_r2d.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
for k in reach_groups classify_items; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_r2d_$k \
      python profiles/profile_pose.py 4 > gpurun_out/ncu_r2d_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:box_tiles_warp -s 6 -c 2 -f -o gpurun_out/prof_r2d_box_tiles_warp \
    python profiles/profile_pose.py 4 > gpurun_out/ncu_r2d_box_tiles_warp.log 2>&1
ncu --section SpeedOfLight --section LaunchStats --section Occupancy --clock-control none -k regex:'conv|maxpool|head' -s 20 -c 12 -f \
    -o gpurun_out/prof_r2d_cnn_trunk python profiles/cnn_time.py > gpurun_out/ncu_r2d_cnn.log 2>&1
for tool in memcheck racecheck synccheck; do
  echo "== $tool" >> gpurun_out/sanitizer_r2d.txt
  timeout 600 compute-sanitizer --tool $tool python profiles/debug_reach.py 20000 2>&1 | grep -E "ERROR SUMMARY|mismatches|RACECHECK SUMMARY" >> gpurun_out/sanitizer_r2d.txt
done
ls -la gpurun_out | tail -8
