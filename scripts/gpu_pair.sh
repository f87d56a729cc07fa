#!/bin/bash
# CTA-pair node kernel (DIFFPHAR_NODE_PAIR=1): hang-guarded smoke, parity subset, same-box bench A/B, node timeline.
TAG=${1:-pair}
OUT=gpurun_out; mkdir -p $OUT
export DIFFPHAR_NODE_PAIR=1
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1 || { echo "pair smoke failed/hung"; tail -15 $OUT/${TAG}_smoke.txt; exit 1; }
tail -3 $OUT/${TAG}_smoke.txt
timeout 500 python -m pytest tests -m gpu -x -q -k "${KEXPR:-tensor_core or sample_given or graph_replay or full_size or large_pocket}" > $OUT/${TAG}_pytest.txt 2>&1
tail -6 $OUT/${TAG}_pytest.txt
for P in 0 1 0 1; do
  DIFFPHAR_NODE_PAIR=$P timeout 200 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_$P.json 2> $OUT/${TAG}_bench_$P.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_$P.json")); r = d["roofline"]
    print("[pair=$P] samples/s", round(d["value"], 1), "step_us", round(d["denoise_step_us"], 1), {k: round(v, 2) for k, v in r["kernel_ms_by_kind"].items()})
except Exception as e:
    print("bench pair=$P failed:", e); print(open("$OUT/${TAG}_bench_$P.err").read()[-1500:])
PY
done
