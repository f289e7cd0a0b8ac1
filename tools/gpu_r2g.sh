#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gaps_gpu.py -q -m gpu -k "attention" 2>&1 | tail -40 > gpurun_out/r2g_attn.log; echo "exit $?" >> gpurun_out/r2g_attn.log )
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2g_attn.log | head -30
echo "--- short kernel"; B=24 timeout 120 python tools/bench_attn.py 2>&1 | tail -8
echo "--- flash kernel (A2V_ATTN_SHORT=0)"; A2V_ATTN_SHORT=0 B=24 timeout 120 python tools/bench_attn.py 2>&1 | grep fwd
