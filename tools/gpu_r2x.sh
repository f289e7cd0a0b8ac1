#!/bin/bash
mkdir -p gpurun_out
for sk in 0 1; do for mode in 1 0; do echo "long=$mode skip=$sk"; SKIP_FAR=$sk A2V_ATTN_LONG=$mode ONLY=teacher B=24 timeout 60 python tools/bench_attn.py 2>&1 | grep -E "teacher|rror"; done; done
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention" 2>&1 | tail -15 > gpurun_out/r2x_attn.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2x_attn.log | head -20
A2V_ATTN_LONG=1 ONLY=teacher B=24 timeout 120 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_long --launch-skip 3 --launch-count 2 -f -o /tmp/attn_x python tools/bench_attn.py > gpurun_out/r2x_ncu.log 2>&1
ncu -i /tmp/attn_x.ncu-rep --page raw --csv > gpurun_out/r2x_attn.raw.csv 2>/dev/null
ncu -i /tmp/attn_x.ncu-rep --page source --csv > gpurun_out/r2x_attn.source.csv 2>/dev/null
gzip -f gpurun_out/r2x_attn.source.csv
