#!/bin/bash
# round-2 session-2 call: full GPU tests with the split node tiling, then same-box A/B of the node split and the NOINIT edge build, node timelines
TAG=${1:-r06b}
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1 || { echo "smoke failed/hung"; tail -8 $OUT/${TAG}_smoke.txt; exit 1; }
tail -3 $OUT/${TAG}_smoke.txt
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1; tail -3 $OUT/${TAG}_pytest.txt
for W in config2 config3; do
for V in "$@"; do
  NAME=${V%%:*}; ENVV=${V#*:}
  env $ENVV timeout 400 python bench.py --workload $W --steps 3 --warmup 3 --no-also --no-cpu-baseline > $OUT/${TAG}_ab_${NAME}_${W}.json 2> $OUT/${TAG}_ab_${NAME}_${W}.err
  python - <<PY
import json
try:
    j = json.load(open("$OUT/${TAG}_ab_${NAME}_${W}.json"))
    print("$NAME", "$W", "samples/s %.2f" % j["value"], "step_us %.1f" % j["denoise_step_us"], "msg_us %.2f" % j["roofline"]["avg_launch_us"], "frac %.3f" % j["roofline"]["frac"],
          "node_us %.2f" % j["roofline_node"]["avg_launch_us"], "node_frac %.3f" % j["roofline_node"]["frac"], {k: round(v, 3) for k, v in j["roofline"]["kernel_ms_by_kind"].items()})
except Exception as e:
    print("$NAME failed", e); print(open("$OUT/${TAG}_ab_${NAME}_${W}.err").read()[-800:])
PY
done
done | tee $OUT/${TAG}_ab_summary.txt
for CTA in 0 100; do
  timeout 120 python scripts/node_trace.py f16fast $CTA 3 > $OUT/${TAG}_node_timeline_cta${CTA}.txt 2>&1; head -3 $OUT/${TAG}_node_timeline_cta${CTA}.txt | cut -c1-400
done
