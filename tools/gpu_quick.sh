#!/bin/bash
# Short GPU-box visit: parity tests, bench line without the CPU baseline, in-situ per-op profile.
mkdir -p gpurun_out
B=${B:-24}
( timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --batch $B > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1800 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
B=$B timeout 300 python tools/profile_step.py > gpurun_out/profile_step.txt 2>&1; head -${TOPN:-50} gpurun_out/profile_step.txt
