#!/bin/bash
# bench lines only (smoke first as a hang guard): bash scripts/gpu_bench_quick.sh <tag> [workloads]
TAG=${1:-bq}; WL=${2:-"config2 config3"}
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1 || { echo "smoke failed/hung"; tail -8 $OUT/${TAG}_smoke.txt; exit 1; }
for W in $WL; do
  timeout 400 python bench.py --workload $W --steps 3 --warmup 3 --no-also --no-cpu-baseline > $OUT/${TAG}_bench_${W}.json 2> $OUT/${TAG}_bench_${W}.err
  python - <<PY
import json
try:
    j = json.load(open("$OUT/${TAG}_bench_${W}.json")); r = j["roofline"]; n = j["roofline_node"]
    print("$W samples/s %.2f step_us %.1f | msg in-graph %.2f us frac %.3f (eager %.2f us frac %.3f) share %.3f | node in-graph %.2f us frac %.3f (eager %.2f) share %.3f" % (
        j["value"], j["denoise_step_us"], r["avg_launch_us"], r["frac"], r["avg_launch_us_eager"], r["frac_eager"], r["share_of_step"],
        n["avg_launch_us"], n["frac"], n["avg_launch_us_eager"], n["share_of_step"]))
    print("  back-to-back msg", r.get("back_to_back"))
    print("  in-graph", {k: round(v, 3) for k, v in r["kernel_ms_by_kind"].items()}, "calls", r["denoiser_calls_timed"])
    print("  eager   ", {k: round(v, 3) for k, v in r["kernel_ms_by_kind_eager"].items()})
    print("  rooflines", [(e["kernel"][:12], round(e["us_per_denoiser_call"], 1), e["launches_per_call"]) for e in j["rooflines"]])
except Exception as e:
    print("$W failed", e); print(open("$OUT/${TAG}_bench_${W}.err").read()[-1500:])
PY
done
