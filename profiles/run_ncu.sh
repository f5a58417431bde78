#!/bin/bash
# Run under gpurun (1 GPU). Produces the launch list and one full capture of the dominant kernel.
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:check_items_warp -s 3 -c 1 -f -o gpurun_out/prof_warp \
    python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
