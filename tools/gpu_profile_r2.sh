#!/bin/bash
# Round-2 evidence: ncu launch list of the bench command, --set full captures of the kernel families VERDICT r1 asked for
# (dominant GEMM with DRAM traffic, slab convs, attention forward/backward, every memory-bound kernel). Only the CSV
# exports travel back (gpurun_out/ is capped at 64 MiB): the .ncu-rep files are deleted on the box.
mkdir -p gpurun_out
B=${B:-24}
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batch $B > gpurun_out/r2_ncu_bench.log 2>&1
python tools/launch_shares.py gpurun_out/r2_launches.csv 60 > gpurun_out/r2_launch_shares.md 2>&1; head -12 gpurun_out/r2_launch_shares.md
gzip -f gpurun_out/r2_launches.csv
cap() {  # name regex skip count extra-flags
  timeout 600 ncu --set full --clock-control none $5 -k regex:"$2" --launch-skip $3 --launch-count $4 --kill on -f \
    -o /tmp/r2_full_$1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch $B > gpurun_out/r2_full_$1.log 2>&1
  tail -1 gpurun_out/r2_full_$1.log | cut -c1-160
  ncu -i /tmp/r2_full_$1.ncu-rep --page raw --csv > gpurun_out/r2_full_$1.raw.csv 2>/dev/null
  if [ -n "$5" ]; then ncu -i /tmp/r2_full_$1.ncu-rep --page source --csv > gpurun_out/r2_full_$1.source.csv 2>/dev/null; gzip -f gpurun_out/r2_full_$1.source.csv; fi
  rm -f /tmp/r2_full_$1.ncu-rep
}
# one priming fwd/bwd + 3 warm-up steps precede the profiled window: skip counts are per kernel-name match
cap gemm "gemm2cta" 900 12 "--import-source on"
cap slab "conv_slab" 100 6 ""
cap attn "attn_fwd|attn_bwd" 200 8 "--import-source on"
cap mem "d2v_loss|target_|ema_step|row_gather|resln|rowln_gelu|adamw|clone_sum|colsum|dgelu|mask_index|relayout|sumsq" 1400 36 ""
du -sh gpurun_out; ls -la gpurun_out | head -40
