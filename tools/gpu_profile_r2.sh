#!/bin/bash
# Round-2 evidence: ncu launch list of the bench command, --set full captures of the kernel families VERDICT r1 asked for
# (dominant GEMM with DRAM traffic, slab convs, attention forward/backward, every memory-bound kernel).
mkdir -p gpurun_out
B=${B:-24}
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batch $B > gpurun_out/r2_ncu_bench.log 2>&1
python tools/launch_shares.py gpurun_out/r2_launches.csv 60 > gpurun_out/r2_launch_shares.md 2>&1; head -12 gpurun_out/r2_launch_shares.md
cap() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" --launch-skip $3 --launch-count $4 --kill on -f \
    -o gpurun_out/r2_full_$1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch $B > gpurun_out/r2_full_$1.log 2>&1
  tail -1 gpurun_out/r2_full_$1.log | cut -c1-200
}
# one priming fwd/bwd + 3 warm-up steps precede the profiled window: skip counts are per kernel-name match
cap gemm "gemm2cta" 900 12
cap slab "conv_slab" 100 6
cap attn "attn_fwd|attn_bwd" 200 8
cap mem "d2v_loss|target_|ema_step|row_gather|resln|rowln_gelu|adamw|clone_sum|colsum|dgelu|mask_index|relayout|sumsq" 1400 60
ls -la gpurun_out/r2_full_*.ncu-rep
