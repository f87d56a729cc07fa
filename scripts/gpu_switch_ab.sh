#!/bin/bash
# A/B of one environment switch that changes kernel code paths: hang-guarded smoke and a parity subset WITH the switch,
# then same-box bench lines without / with it, and the edge timeline with it.
#   gpurun --timeout 900 -- 'bash scripts/gpu_switch_ab.sh <tag> "<VAR=val ...>" [workloads]'
TAG=${1:-sw}; SETTING=$2; WL=${3:-config2}
OUT=gpurun_out; mkdir -p $OUT
env $SETTING timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1 || { echo "smoke with [$SETTING] failed/hung"; tail -8 $OUT/${TAG}_smoke.txt; exit 1; }
tail -2 $OUT/${TAG}_smoke.txt
env $SETTING timeout 400 python -m pytest tests -m gpu -x -q -k "${KEXPR:-tensor_core or graph_replay or segmented or sample_given}" > $OUT/${TAG}_pytest.txt 2>&1
tail -3 $OUT/${TAG}_pytest.txt
for rep in 1 2; do
  for S in "A=0" "$SETTING"; do
    for W in $WL; do
      env $S timeout 200 python bench.py --workload $W --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
      python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench.json")); r = d["roofline"]
    print("[$S] $W samples/s", round(d["value"], 1), "step_us", round(d["denoise_step_us"], 1), "edge_us", round(r["avg_launch_us"], 2), "frac", round(r["frac"], 3),
          {k: round(v, 2) for k, v in r["kernel_ms_by_kind"].items()})
except Exception as e:
    print("bench [$S] $W failed:", e); print(open("$OUT/${TAG}_bench.err").read()[-1200:])
PY
    done
  done
done
env $SETTING timeout 120 python scripts/edge_trace.py f16fast > $OUT/${TAG}_edge_timeline.txt 2>&1
head -8 $OUT/${TAG}_edge_timeline.txt
