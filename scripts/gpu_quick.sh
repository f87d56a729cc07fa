#!/bin/bash
# Quick GPU iteration: tensor-core parity subset, bench line (no CPU baseline), clock64 timelines.
#   gpurun --timeout 900 -- 'bash scripts/gpu_quick.sh <tag> [pytest -k expression]'
TAG=${1:-q}
KEXPR=${2:-"tensor_core or sample_given or graph_replay or full_size"}
OUT=gpurun_out
mkdir -p $OUT
# a deadlocked kernel must not eat the GPU budget: smoke first under a short timeout, bail out if it hangs
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1 || { echo "smoke failed/hung"; tail -5 $OUT/${TAG}_smoke.txt; exit 1; }
tail -2 $OUT/${TAG}_smoke.txt
timeout 400 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $OUT/${TAG}_pytest.txt 2>&1
tail -4 $OUT/${TAG}_pytest.txt
timeout 200 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench.json"))
    r = d["roofline"]
    print("samples/s", round(d["value"], 1), "step_us", round(d["denoise_step_us"], 1), "edge_us", round(r["avg_launch_us"], 2),
          "frac", round(r["frac"], 3), {k: round(v, 2) for k, v in r["kernel_ms_by_kind"].items()})
except Exception as e:
    print("bench failed:", e); print(open("$OUT/${TAG}_bench.err").read()[-2000:])
PY
timeout 120 python scripts/edge_trace.py bf16 > $OUT/${TAG}_edge_timeline.txt 2>&1
timeout 120 python scripts/node_trace.py bf16 > $OUT/${TAG}_node_timeline.txt 2>&1
