#!/bin/bash
# Node kernel with the multicast weight stream (DIFFPHAR_NODE_MC): hang-guarded smoke, the GPU tests, same-box A/B, timeline.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_mc_ab.sh <tag>'
TAG=${1:-mc}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $OUT/${TAG}_smi.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1 || { echo "smoke failed/hung"; tail -8 $OUT/${TAG}_smoke.txt; exit 1; }
tail -3 $OUT/${TAG}_smoke.txt
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1; tail -3 $OUT/${TAG}_pytest.txt
for V in "base:DIFFPHAR_NODE_MC=0" "mc:DIFFPHAR_NODE_MC=1" "base2:DIFFPHAR_NODE_MC=0" "mc2:DIFFPHAR_NODE_MC=1"; do
  for W in ${WORKLOADS:-config2 config3}; do
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
timeout 120 python scripts/node_trace.py f16fast > $OUT/${TAG}_node_timeline.txt 2>&1; head -30 $OUT/${TAG}_node_timeline.txt
