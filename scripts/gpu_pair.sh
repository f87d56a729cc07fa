#!/bin/bash
# CTA-pair node kernel (DIFFPHAR_NODE_PAIR=1) experiments: hang-guarded smoke + parity, then settings given as "VAR=val ..." strings; same-box bench + node timeline per setting.
TAG=${1:-pair}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 180 env DIFFPHAR_NODE_PAIR=1 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1 || { echo "pair smoke failed/hung"; tail -15 $OUT/${TAG}_smoke.txt; exit 1; }
tail -2 $OUT/${TAG}_smoke.txt
timeout 300 env DIFFPHAR_NODE_PAIR=1 python -m pytest tests -m gpu -x -q -k "${KEXPR:-tensor_core or graph_replay}" > $OUT/${TAG}_pytest.txt 2>&1
tail -3 $OUT/${TAG}_pytest.txt
i=0
for SETTING in "$@"; do
  env $SETTING timeout 200 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_$i.json 2> $OUT/${TAG}_bench_$i.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_$i.json")); r = d["roofline"]
    print("[$SETTING] samples/s", round(d["value"], 1), "step_us", round(d["denoise_step_us"], 1), {k: round(v, 2) for k, v in r["kernel_ms_by_kind"].items()})
except Exception as e:
    print("bench [$SETTING] failed:", e); print(open("$OUT/${TAG}_bench_$i.err").read()[-1500:])
PY
  env $SETTING timeout 120 python scripts/node_trace.py f16fast 2>&1 | head -1 | tee $OUT/${TAG}_node_timeline_$i.txt
  i=$((i+1))
done
