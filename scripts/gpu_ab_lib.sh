#!/bin/bash
# same-box A/B of library builds / env switches at one or more precisions; smoke (hang guard) + the tensor-core parity subset first
#   bash scripts/gpu_ab_lib.sh <tag> "<precisions>" "name:ENV=.." ...
TAG=$1; PRECS=$2; shift; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1 || { echo "smoke failed/hung"; tail -8 $OUT/${TAG}_smoke.txt; exit 1; }
tail -2 $OUT/${TAG}_smoke.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "${KEXPR:-tensor_core or large or sample_given or segmented or joint}" > $OUT/${TAG}_pytest.txt 2>&1; tail -2 $OUT/${TAG}_pytest.txt
for P in $PRECS; do
for W in ${WORKLOADS:-config2 config3}; do
for V in "$@"; do
  NAME=${V%%:*}; ENVV=${V#*:}
  env $ENVV timeout 400 python bench.py --precision $P --workload $W --steps 3 --warmup 3 --no-also --no-cpu-baseline > $OUT/${TAG}_ab_${NAME}_${W}_${P}.json 2> $OUT/${TAG}_ab_${NAME}_${W}_${P}.err
  python - <<PY
import json
try:
    j = json.load(open("$OUT/${TAG}_ab_${NAME}_${W}_${P}.json"))
    print("$NAME", "$W", "$P", "samples/s %.2f" % j["value"], "step_us %.1f" % j["denoise_step_us"], "msg_us %.2f" % j["roofline"]["avg_launch_us"], "frac %.3f" % j["roofline"]["frac"],
          "node_us %.2f" % j["roofline_node"]["avg_launch_us"], {k: round(v, 3) for k, v in j["roofline"]["kernel_ms_by_kind"].items()})
except Exception as e:
    print("$NAME failed", e); print(open("$OUT/${TAG}_ab_${NAME}_${W}_${P}.err").read()[-800:])
PY
done
done
done | tee $OUT/${TAG}_ab_summary.txt
