#!/bin/bash
# A/B of environment switches on one box: smoke (hang guard) then bench lines per setting.
#   gpurun --timeout 900 -- 'bash scripts/gpu_env_ab.sh <tag> "<VAR=val ...>" "<VAR=val ...>" ...'
TAG=${1:-env}; shift
OUT=gpurun_out; mkdir -p $OUT
PREC=${PREC:-f16fast}
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1 || { echo "smoke failed/hung"; tail -5 $OUT/${TAG}_smoke.txt; exit 1; }
tail -2 $OUT/${TAG}_smoke.txt
i=0
for SETTING in "$@"; do
  for W in ${WORKLOADS:-config2}; do
    env $SETTING timeout 200 python bench.py --precision $PREC --workload $W --no-cpu-baseline > $OUT/${TAG}_bench_${i}_${W}.json 2> $OUT/${TAG}_bench_${i}_${W}.err
    python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_${i}_${W}.json"))
    r = d["roofline"]
    print("[$SETTING] $W samples/s", round(d["value"], 1), "step_us", round(d["denoise_step_us"], 1), "e2e", round(d["e2e"]["value"], 1), "edge_us", round(r["avg_launch_us"], 2),
          "frac", round(r["frac"], 3), {k: round(v, 2) for k, v in r["kernel_ms_by_kind"].items()})
except Exception as e:
    print("bench [$SETTING] $W failed:", e); print(open("$OUT/${TAG}_bench_${i}_${W}.err").read()[-1500:])
PY
  done
  i=$((i+1))
done
if [ -n "${KEXPR:-}" ]; then
  timeout 500 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $OUT/${TAG}_pytest.txt 2>&1; tail -3 $OUT/${TAG}_pytest.txt
fi
if [ "${TRACE:-0}" = "1" ]; then
  timeout 120 python scripts/edge_trace.py $PREC > $OUT/${TAG}_edge_timeline.txt 2>&1
  timeout 120 python scripts/node_trace.py $PREC > $OUT/${TAG}_node_timeline.txt 2>&1
fi
