#!/bin/bash
# r02 final code: launch list of the default bench command + full captures of the pose-path kernels (bench workload and
# the rough level) + a light capture of every motion-cost trunk kernel
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 120 --csv --log-file gpurun_out/launches_r2b.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
for k in classify_items box_tiles_warp; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 2 -f -o gpurun_out/prof_r2b_$k \
      python profiles/profile_pose.py 4 > gpurun_out/ncu_r2b_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:box_tiles_warp -s 4 -c 2 -f -o gpurun_out/prof_r2b_rough_box_tiles_warp \
    python profiles/profile_pose.py 4 rough > gpurun_out/ncu_r2b_rough.log 2>&1
ncu --section SpeedOfLight --section LaunchStats --section Occupancy --clock-control none -k regex:'conv|maxpool|head' -s 20 -c 12 -f \
    -o gpurun_out/prof_r2b_cnn_trunk python profiles/cnn_time.py > gpurun_out/ncu_r2b_cnn.log 2>&1
ls -la gpurun_out | tail -6
