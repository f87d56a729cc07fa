#!/bin/bash
# same-box A/B of library variants / environment switches: bash scripts/gpu_ab_only.sh <tag> <workload> "name:ENV=.." ...
set -u
TAG=$1; W=$2; shift; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1 || { echo "smoke failed"; tail -5 $OUT/${TAG}_smoke.txt; }
for V in "$@"; do
  NAME=${V%%:*}; ENVV=${V#*:}
  env $ENVV timeout 400 python bench.py --workload $W --steps ${STEPS:-5} --warmup 3 --no-also --no-cpu-baseline > $OUT/${TAG}_ab_${NAME}.json 2> $OUT/${TAG}_ab_${NAME}.err
  python - <<PY
import json
try:
    j = json.load(open("$OUT/${TAG}_ab_${NAME}.json"))
    print("$NAME", "$W", "samples/s %.2f" % j["value"], "step_us %.1f" % j["denoise_step_us"], "msg_us %.2f" % j["roofline"]["avg_launch_us"], "frac %.3f" % j["roofline"]["frac"], {k: round(v, 3) for k, v in j["roofline"]["kernel_ms_by_kind"].items()})
except Exception as e:
    print("$NAME failed", e); print(open("$OUT/${TAG}_ab_${NAME}.err").read()[-800:])
PY
done | tee $OUT/${TAG}_ab_summary.txt
