#!/bin/bash
# In-graph cost of each kernel family: replay the captured step with one family skipped (DIFFPHAR_SKIP bits:
# 1 edge message, 2 node, 4 coordinate edge, 8 coordinate finish, 16 radius graph, 32 encode + decode, 64 none) and
# difference the step time against the full run.  Results of the skipped runs are garbage by construction.
OUT=gpurun_out; TAG=${1:-bd}; mkdir -p $OUT
for M in 0 1 2 4 8 12 16 32 0; do
  DIFFPHAR_SKIP=$M timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('skip=$M step_us', round(d['denoise_step_us'],1))" 2>&1 | tee -a $OUT/${TAG}_breakdown.txt
done
