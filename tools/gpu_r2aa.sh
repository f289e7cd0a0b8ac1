#!/bin/bash
for sp in 0 1 2 3; do echo "stream spin=$sp"; A2V_ATTN_SPIN=$sp SKIP_FAR=1 ONLY=teacher B=24 timeout 60 python tools/bench_attn.py 2>&1 | grep -E "teacher|rror"; done
