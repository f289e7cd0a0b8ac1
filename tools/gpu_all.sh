#!/bin/bash
# full GPU suite (as the driver runs it) + bench line without the CPU baseline
mkdir -p gpurun_out
TAG=${TAG:-r2}
( timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -80 > gpurun_out/${TAG}_pytest_gpu.log; echo "exit $?" >> gpurun_out/${TAG}_pytest_gpu.log )
grep -E "passed|failed|FAILED|^E  " gpurun_out/${TAG}_pytest_gpu.log | head -40
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench.json'));print('bench', d['value'], d['ms_per_step'], d['step_tensor_frac'], d['e2e']['value'])"; tail -3 gpurun_out/${TAG}_bench.err
