#!/bin/bash
mkdir -p gpurun_out
echo "stream skip=1"; SKIP_FAR=1 ONLY=teacher B=24 timeout 60 python tools/bench_attn.py 2>&1 | grep -E "teacher|rror"
echo "flash skip=1"; A2V_ATTN_STREAM=0 SKIP_FAR=1 ONLY=teacher B=24 timeout 60 python tools/bench_attn.py 2>&1 | grep -E "teacher|rror"
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" 2>&1 | tail -25 > gpurun_out/r2z_attn.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2z_attn.log | head -20
SKIP_FAR=1 ONLY=teacher B=24 timeout 120 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_stream --launch-skip 3 --launch-count 2 -f -o /tmp/attn_z python tools/bench_attn.py > gpurun_out/r2z_ncu.log 2>&1
ncu -i /tmp/attn_z.ncu-rep --page raw --csv > gpurun_out/r2z_attn.raw.csv 2>/dev/null
ncu -i /tmp/attn_z.ncu-rep --page source --csv > gpurun_out/r2z_attn.source.csv 2>/dev/null
gzip -f gpurun_out/r2z_attn.source.csv
