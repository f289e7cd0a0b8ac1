#!/bin/bash
# One GPU-box visit: parity tests, default bench line (with the CPU baseline), in-situ per-op profile, ncu launch list of
# the bench command, ncu --set full of the hot kernels (one launch each of the dominant shapes).
mkdir -p gpurun_out
B=${B:-24}
( timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json
B=$B timeout 200 python tools/profile_step.py > gpurun_out/profile_step.txt 2>&1; head -12 gpurun_out/profile_step.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batch $B > gpurun_out/ncu_bench.log 2>&1
python tools/launch_shares.py gpurun_out/launches.csv 50 > gpurun_out/launch_shares.md 2>&1; head -8 gpurun_out/launch_shares.md
# full capture: skip the priming + first warm-up step, then 1 launch in 7 of the GEMM / attention / conv / LN kernels
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"gemm2cta|gemm_tcgen05|attn_fwd|attn_bwd|slab|resln|rowln_gelu" \
  --launch-skip 1500 --launch-count 14 --kill on -f -o gpurun_out/r1_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch 16 \
  > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out | head -30
