#!/bin/bash
# round-2 first visit: the new parity tests first (no -x: every failure is listed), then the old suite, then a bench line
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_parity_gaps_gpu.py tests/test_dropin_gpu.py tests/test_trainer_gpu.py -q -m gpu 2>&1 | tail -80 > gpurun_out/r2a_new_tests.log; echo "exit $?" >> gpurun_out/r2a_new_tests.log )
tail -40 gpurun_out/r2a_new_tests.log
( timeout 600 python -m pytest tests -q -m gpu -x --deselect tests/test_parity_gaps_gpu.py --deselect tests/test_dropin_gpu.py --deselect tests/test_trainer_gpu.py 2>&1 | tail -15 > gpurun_out/r2a_old_tests.log )
tail -5 gpurun_out/r2a_old_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json
