#!/bin/bash
# r02 v5 (group kernel on): full captures of the three pose-path box kernels on the bench workload
mkdir -p gpurun_out
for k in reach_groups classify_items; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_r2c_$k \
      python profiles/profile_pose.py 4 > gpurun_out/ncu_r2c_$k.log 2>&1
done
ls -la gpurun_out | tail -4
