#!/bin/bash
# ncu launch list of the bench command at the final state of the round (batch 32)
mkdir -p gpurun_out
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2d_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 32 > gpurun_out/r2d_ncu_bench.log 2>&1
python tools/launch_shares.py gpurun_out/r2d_launches.csv 70 > gpurun_out/r2d_launch_shares.md 2>&1; head -16 gpurun_out/r2d_launch_shares.md
gzip -f gpurun_out/r2d_launches.csv
