#!/bin/bash
# A/B of the edge kernel's arithmetic modes on one box: smoke (hang guard), tensor-core parity subset, actual errors,
# bench lines per mode (config 2 and config 3), clock64 timeline, one ncu --set full capture at config 3.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_ab.sh <tag> "<modes>"'
TAG=${1:-ab}
MODES=${2:-"bf16 f16fast f16fast32"}
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1 || { echo "smoke failed/hung"; tail -5 $OUT/${TAG}_smoke.txt; exit 1; }
tail -3 $OUT/${TAG}_smoke.txt
timeout 500 python -m pytest tests -m gpu -x -q -k "${KEXPR:-tensor_core or sample_given or graph_replay or full_size or large_pocket}" > $OUT/${TAG}_pytest.txt 2>&1
tail -4 $OUT/${TAG}_pytest.txt
timeout 200 python scripts/tc_error.py > $OUT/${TAG}_tc_error.txt 2>&1
head -30 $OUT/${TAG}_tc_error.txt
for M in $MODES; do
  for W in config2 config3; do
    timeout 200 python bench.py --precision $M --workload $W --no-cpu-baseline > $OUT/${TAG}_bench_${W}_${M}.json 2> $OUT/${TAG}_bench_${W}_${M}.err
    python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_${W}_${M}.json"))
    r = d["roofline"]
    print("$W $M samples/s", round(d["value"], 1), "step_us", round(d["denoise_step_us"], 1), "edge_us", round(r["avg_launch_us"], 2),
          "frac", round(r["frac"], 3), {k: round(v, 2) for k, v in r["kernel_ms_by_kind"].items()})
except Exception as e:
    print("bench $W $M failed:", e); print(open("$OUT/${TAG}_bench_${W}_${M}.err").read()[-1500:])
PY
  done
done
LAST=$(echo $MODES | awk '{print $NF}')
NCU_MODE=${NCU_MODE:-f16fast}
timeout 120 python scripts/edge_trace.py $NCU_MODE > $OUT/${TAG}_edge_timeline_${NCU_MODE}.txt 2>&1
if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:edge_tc_kernel -s 12 -c 1 -o $OUT/${TAG}_config3_${NCU_MODE}_edge -f \
      python bench.py --precision $NCU_MODE --workload config3 --steps 1 --warmup 1 --timesteps 4 --no-cpu-baseline > $OUT/${TAG}_ncu_config3.log 2>&1
  tail -2 $OUT/${TAG}_ncu_config3.log
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:edge_tc_kernel -s 32 -c 1 -o $OUT/${TAG}_config2_${NCU_MODE}_edge -f \
      python bench.py --precision $NCU_MODE --workload config2 --steps 1 --warmup 1 --timesteps 8 --no-cpu-baseline > $OUT/${TAG}_ncu_config2.log 2>&1
  tail -2 $OUT/${TAG}_ncu_config2.log
fi
ls -la $OUT | tail -30
