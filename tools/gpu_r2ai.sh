#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu -k "pair" 2>&1 | tail -12 > gpurun_out/r2ai_pair.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2ai_pair.log | head
for f in 1 0; do echo "fuse colsum=$f"; A2V_FUSE_COLSUM=$f B=24 TOP=70 timeout 300 python tools/profile_step.py > gpurun_out/r2ai_profile_step_$f.txt 2>&1; grep -E "step |dgelu|epi=0001|a2v_colsum" gpurun_out/r2ai_profile_step_$f.txt; done
