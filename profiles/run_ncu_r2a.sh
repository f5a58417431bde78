#!/bin/bash
# r02 v1: full captures of the pose-path kernels after the thread-level reach-box stages went in (bench workload + rough)
mkdir -p gpurun_out
for k in classify_items box_tiles_warp; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 2 -f -o gpurun_out/prof_r2a_$k \
      python profiles/profile_pose.py 4 > gpurun_out/ncu_r2a_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:box_tiles_warp -s 4 -c 2 -f -o gpurun_out/prof_r2a_rough_box_tiles_warp \
    python profiles/profile_pose.py 4 rough > gpurun_out/ncu_r2a_rough.log 2>&1
ls -la gpurun_out | tail -8
