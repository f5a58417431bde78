#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:check_items_warp -s 3 -c 1 -f -o gpurun_out/prof_warp_v2 \
    python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
