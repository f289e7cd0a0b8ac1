#!/bin/bash
# second evidence pass: --set full of the once-per-step memory-bound kernels (loss, targets, EMA, AdamW, mask index,
# clone sums, norms, casts) and of the attention forwards, at batch 16 (the per-step kernels' sizes scale with the batch)
mkdir -p gpurun_out
B=${B:-16}
cap() {  # name regex skip count extra-flags
  timeout 400 ncu --set full --clock-control none $5 -k regex:"$2" --launch-skip $3 --launch-count $4 --kill on -f \
    -o /tmp/r2b_full_$1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch $B > gpurun_out/r2b_full_$1.log 2>&1
  tail -1 gpurun_out/r2b_full_$1.log | cut -c1-160
  ncu -i /tmp/r2b_full_$1.ncu-rep --page raw --csv > gpurun_out/r2b_full_$1.raw.csv 2>/dev/null
  if [ -n "$5" ]; then ncu -i /tmp/r2b_full_$1.ncu-rep --page source --csv > gpurun_out/r2b_full_$1.source.csv 2>/dev/null; gzip -f gpurun_out/r2b_full_$1.source.csv; fi
  rm -f /tmp/r2b_full_$1.ncu-rep
}
cap step "d2v_loss|target_|ema_kernel|adamw|mask_index|clone_sum|sumsq|clip_coef|cast_|sinc_|mixup_|rowln128|relayout_batch|neigh_index" 40 40 ""
cap attnf "attn_fwd_stream|attn_fwd_short|attn_qk_bound|attn_fwd_tcgen05" 150 9 "--import-source on"
cap rows "resln_fwd|rowln_gelu_fwd|rowln_fwd|rowln_bwd" 300 12 ""
ls -la gpurun_out | grep r2b
