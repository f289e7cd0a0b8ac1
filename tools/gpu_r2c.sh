#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_finetune_gpu.py -q -m gpu 2>&1 | tail -120 > gpurun_out/r2c_finetune.log; echo "exit $?" >> gpurun_out/r2c_finetune.log )
tail -70 gpurun_out/r2c_finetune.log
( timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_dropin_gpu.py tests/test_kernels_gpu.py -q -m gpu -x -k "not alibi_locality" 2>&1 | tail -15 > gpurun_out/r2c_regress.log )
tail -6 gpurun_out/r2c_regress.log
