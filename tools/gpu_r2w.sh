#!/bin/bash
mkdir -p gpurun_out
for mode in 1 0; do
  K=$([ $mode = 1 ] && echo attn_fwd_long || echo attn_fwd_tcgen05)
  A2V_ATTN_LONG=$mode ONLY=teacher B=24 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 3 --launch-count 2 -f -o /tmp/attn_$mode python tools/bench_attn.py > gpurun_out/r2w_ncu_$mode.log 2>&1
  ncu -i /tmp/attn_$mode.ncu-rep --page raw --csv > gpurun_out/r2w_attn_$mode.raw.csv 2>/dev/null
  ncu -i /tmp/attn_$mode.ncu-rep --page source --csv > gpurun_out/r2w_attn_$mode.source.csv 2>/dev/null
  gzip -f gpurun_out/r2w_attn_$mode.source.csv
done
for sk in 0 1; do for mode in 1 0; do echo "long=$mode skip=$sk"; SKIP_FAR=$sk A2V_ATTN_LONG=$mode ONLY=teacher B=24 timeout 120 python tools/bench_attn.py 2>&1 | grep teacher; done; done
ls -la gpurun_out | grep r2w
