#!/bin/bash
mkdir -p gpurun_out
python tools/trace_attn_short.py 2>&1 | head -32
( timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gaps_gpu.py -q -m gpu -k "attention" 2>&1 | tail -40 > gpurun_out/r2k_attn.log; echo "exit $?" >> gpurun_out/r2k_attn.log )
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2k_attn.log | head -30
B=24 timeout 120 python tools/bench_attn.py 2>&1 | grep fwd
