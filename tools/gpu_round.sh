#!/bin/bash
# One GPU-box visit: parity tests, default bench line, in-situ per-op profile, ncu launch list, ncu --set full of the hot kernels.
mkdir -p gpurun_out
B=${B:-24}
( timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json
B=$B timeout 300 python tools/profile_step.py > gpurun_out/profile_step.txt 2>&1; head -60 gpurun_out/profile_step.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --batch $B > gpurun_out/ncu_bench.log 2>&1
python tools/launch_shares.py gpurun_out/launches.csv 45 > gpurun_out/launch_shares.md 2>&1; head -5 gpurun_out/launch_shares.md
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|attn_|slab" --launch-skip 900 -c 14 \
  -f -o gpurun_out/r1_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 8 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
