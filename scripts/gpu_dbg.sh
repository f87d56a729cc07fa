#!/bin/bash
# Timing experiments with DIFFPHAR_DBG bits (results of dbg != 0 runs are wrong by construction: DIFFPHAR_SKIP=64 tells
# bench.py not to assert on them).  Edge kernel: 2 = never reload Pa, 4 = no L1 touch of the next tile's Pa rows.
#   gpurun --timeout 900 -- 'bash scripts/gpu_dbg.sh <tag> "<dbg values>"'
TAG=${1:-dbg}; VALS=${2:-"0 2 4 6 0"}
OUT=gpurun_out; mkdir -p $OUT
export DIFFPHAR_SKIP=64
for D in $VALS; do
  DIFFPHAR_DBG=$D timeout 150 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2>/dev/null
  python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench.json")); r = d["roofline"]
    print("dbg=$D step_us", round(d["denoise_step_us"], 1), "edge_us", round(r["avg_launch_us"], 2), {k: round(v, 2) for k, v in r["kernel_ms_by_kind"].items()})
except Exception as e:
    print("dbg=$D failed:", e)
PY
done | tee $OUT/${TAG}_dbg.txt
for D in 0 2; do echo "producer rows, dbg=$D"; DIFFPHAR_DBG=$D timeout 100 python scripts/edge_trace.py f16fast 2>&1 | grep "producer" | head -5; done | tee -a $OUT/${TAG}_dbg.txt
