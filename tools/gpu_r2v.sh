#!/bin/bash
# new long-sequence attention forward: parity first (bounded), then A/B timing, then the whole suite + bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention" 2>&1 | tail -15 > gpurun_out/r2v_attn.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2v_attn.log | head -20
if grep -q "failed\|FAILED\|error" gpurun_out/r2v_attn.log; then echo "ATTN TESTS FAILED"; cat gpurun_out/r2v_attn.log; fi
echo "--- long kernel"; B=24 timeout 120 python tools/bench_attn.py 2>&1 | grep teacher
echo "--- flash kernel"; A2V_ATTN_LONG=0 B=24 timeout 120 python tools/bench_attn.py 2>&1 | grep teacher
TAG=r2v bash tools/gpu_all.sh
