#!/bin/bash
# final-state evidence: ncu launch list of the bench command (batch 32, one timed step) and --set full of the kernels
# that changed last (fused GELU-backward epilogue of the CTA-pair GEMM, slab convolutions)
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2c_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 32 > gpurun_out/r2c_ncu_bench.log 2>&1
python tools/launch_shares.py gpurun_out/r2c_launches.csv 70 > gpurun_out/r2c_launch_shares.md 2>&1; head -14 gpurun_out/r2c_launch_shares.md
gzip -f gpurun_out/r2c_launches.csv
timeout 400 ncu --set full --clock-control none -k regex:"gemm2cta_kernel<8>|conv_slab_fwd" --launch-skip 60 --launch-count 10 --kill on -f \
  -o /tmp/r2c_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 32 > gpurun_out/r2c_full.log 2>&1
ncu -i /tmp/r2c_full.ncu-rep --page raw --csv > gpurun_out/r2c_full.raw.csv 2>/dev/null
ls -la gpurun_out | grep r2c
