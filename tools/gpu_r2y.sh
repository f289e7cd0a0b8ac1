#!/bin/bash
mkdir -p gpurun_out
for sk in 0 1; do echo "flash skip=$sk"; SKIP_FAR=$sk ONLY=teacher B=24 timeout 60 python tools/bench_attn.py 2>&1 | grep -E "teacher|rror"; done
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention" 2>&1 | tail -25 > gpurun_out/r2y_attn.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2y_attn.log | head -20
ONLY=teacher B=24 timeout 120 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tcgen05 --launch-skip 3 --launch-count 2 -f -o /tmp/attn_y python tools/bench_attn.py > gpurun_out/r2y_ncu.log 2>&1
ncu -i /tmp/attn_y.ncu-rep --page raw --csv > gpurun_out/r2y_attn.raw.csv 2>/dev/null
ncu -i /tmp/attn_y.ncu-rep --page source --csv > gpurun_out/r2y_attn.source.csv 2>/dev/null
gzip -f gpurun_out/r2y_attn.source.csv
