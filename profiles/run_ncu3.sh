#!/bin/bash
# ncu launch list + one full capture of each stage of the pose-validity pass (bench.py workload).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_v3.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
for k in classify_items box_items_warp box_items_block; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_v3_$k \
      python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:conv15_tcgen05 -s 2 -c 1 -f -o gpurun_out/prof_v3_conv15 \
    python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_conv15.log 2>&1
ls -la gpurun_out
