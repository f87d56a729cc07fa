#!/bin/bash
# Full GPU check of the current tree: hang-guarded smoke, the whole -m gpu suite, bench lines (config 2 / 3), timelines.
TAG=${1:-full}
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1 || { echo "smoke failed/hung"; tail -15 $OUT/${TAG}_smoke.txt; exit 1; }
tail -3 $OUT/${TAG}_smoke.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1
tail -8 $OUT/${TAG}_pytest.txt
for W in config2 config3 ${EXTRA_WORKLOADS:-}; do
  timeout 300 python bench.py --workload $W --no-cpu-baseline > $OUT/${TAG}_bench_$W.json 2> $OUT/${TAG}_bench_$W.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_$W.json")); r = d["roofline"]
    print("$W samples/s", round(d["value"], 1), "step_us", round(d["denoise_step_us"], 1), "e2e", round(d["e2e"]["value"], 1), "edge_us", round(r["avg_launch_us"], 2),
          "frac", round(r["frac"], 3), {k: round(v, 2) for k, v in r["kernel_ms_by_kind"].items()})
except Exception as e:
    print("bench $W failed:", e); print(open("$OUT/${TAG}_bench_$W.err").read()[-1500:])
PY
done
timeout 120 python scripts/edge_trace.py f16fast > $OUT/${TAG}_edge_timeline.txt 2>&1
timeout 120 python scripts/node_trace.py f16fast > $OUT/${TAG}_node_timeline.txt 2>&1
head -1 $OUT/${TAG}_node_timeline.txt
