#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_kernels_gpu.py tests/test_parity_gaps_gpu.py tests/test_finetune_gpu.py -q -m gpu -k "48khz or fused or targets_and_loss or large_config_bf16_gradients or dropin_finetune" 2>&1 | tail -60 > gpurun_out/r2d_tests.log; echo "exit $?" >> gpurun_out/r2d_tests.log )
tail -40 gpurun_out/r2d_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 1200 gpurun_out/r2d_bench.json; tail -3 gpurun_out/r2d_bench.err
B=24 timeout 300 python tools/profile_step.py > gpurun_out/r2d_profile_step.txt 2>&1; head -60 gpurun_out/r2d_profile_step.txt
