#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_pose.csv python profiles/profile_pose.py 4 > gpurun_out/pp.log 2>&1
for k in classify_items box_items_warp; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_pose_$k python profiles/profile_pose.py 4 > gpurun_out/ncu_pose_$k.log 2>&1
done
