#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" 2>&1 | tail -60 > gpurun_out/r2b_attn.log; echo "exit $?" >> gpurun_out/r2b_attn.log )
tail -45 gpurun_out/r2b_attn.log
( timeout 600 python -m pytest tests/test_parity_gaps_gpu.py tests/test_trainer_gpu.py -q -m gpu -k "large_config or oracle_optimizer" 2>&1 | tail -30 > gpurun_out/r2b_fix.log )
tail -12 gpurun_out/r2b_fix.log
