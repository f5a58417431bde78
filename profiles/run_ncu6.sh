#!/bin/bash
# r01 v8 (final round-1 code): launch list of the default bench command + full captures of the two dominant pose-path kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 80 --csv --log-file gpurun_out/launches_v8.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
for k in box_items_warp classify_items; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_v8_$k \
      python profiles/profile_pose.py 4 > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out | head -30
