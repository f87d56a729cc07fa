#!/bin/bash
# smoke (bit-exact fp32 sampler against the oracle: the hang / parity guard), then bench lines of the given variants at config 2
#   bash scripts/gpu_ab_quick.sh <tag> "name:ENV=.." ...
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1 || { echo "smoke failed/hung"; tail -8 $OUT/${TAG}_smoke.txt; exit 1; }
tail -3 $OUT/${TAG}_smoke.txt
for V in "$@"; do
  NAME=${V%%:*}; ENVV=${V#*:}
  env $ENVV timeout 400 python bench.py --workload ${W:-config2} --steps 3 --warmup 3 --no-also --no-cpu-baseline > $OUT/${TAG}_ab_${NAME}.json 2> $OUT/${TAG}_ab_${NAME}.err
  python - <<PY
import json
try:
    j = json.load(open("$OUT/${TAG}_ab_${NAME}.json"))
    print("$NAME", "samples/s %.2f" % j["value"], "step_us %.1f" % j["denoise_step_us"], {k: round(v, 3) for k, v in j["roofline"]["kernel_ms_by_kind"].items()})
except Exception as e:
    print("$NAME failed", e); print(open("$OUT/${TAG}_ab_${NAME}.err").read()[-800:])
PY
done | tee $OUT/${TAG}_ab_summary.txt
