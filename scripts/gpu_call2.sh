#!/bin/bash
# Round-2 measurement call: bench line (with side workloads), same-box A/B of the round's two launch fusions,
# precision lines, ncu launch list and --set full captures of the two top kernels at config 2 and of the message kernel at config 3.
set -u
TAG=${1:-r05b}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --sweep ${SWEEP:-16,64} > $OUT/${TAG}_bench_f16fast.json 2> $OUT/${TAG}_bench_f16fast.err
cat $OUT/${TAG}_bench_f16fast.json; tail -3 $OUT/${TAG}_bench_f16fast.err
# same-box A/B (5 runs each): stand-alone coord finish launch; three-launch graph builder
for V in ${AB_SET:-"base:" "scan_kernel_standalone:DIFFPHAR_DBG=128" "base_again:" "scan_kernel_standalone_again:DIFFPHAR_DBG=128" "skip_coord:DIFFPHAR_SKIP=12" "skip_node:DIFFPHAR_SKIP=2"}; do
  NAME=${V%%:*}; ENVV=${V#*:}
  env $ENVV timeout 300 python bench.py --steps 5 --warmup 3 --no-also --no-cpu-baseline > $OUT/${TAG}_ab_${NAME}.json 2> $OUT/${TAG}_ab_${NAME}.err
  python - <<PY
import json
try:
    j = json.load(open("$OUT/${TAG}_ab_${NAME}.json"))
    print("$NAME", "samples/s %.1f" % j["value"], "step_us %.1f" % j["denoise_step_us"], "launches/step %.1f" % (j["gpu_launches"] / j["steps"] / 501.0), {k: round(v, 3) for k, v in j["roofline"]["kernel_ms_by_kind"].items()})
except Exception as e:
    print("$NAME failed", e)
PY
done | tee $OUT/${TAG}_ab_summary.txt
for P in ${EXTRA_PREC:-bf16 f16 tf32 fp32}; do
  timeout 600 python bench.py --precision $P --steps 3 --warmup 3 --no-also --no-cpu-baseline > $OUT/${TAG}_bench_${P}.json 2> $OUT/${TAG}_bench_${P}.err
  python -c "
import json
j = json.load(open('$OUT/${TAG}_bench_${P}.json')); print('$P', 'samples/s %.1f' % j['value'], 'e2e %.1f' % j['e2e']['value'], 'msg frac %.3f' % j['roofline']['frac'])" | tee -a $OUT/${TAG}_precisions.txt
done
if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv \
      --log-file $OUT/${TAG}_f16fast_launches.csv \
      python bench.py --steps 1 --warmup 1 --timesteps 20 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
  python profiles/summarize_launches.py $OUT/${TAG}_f16fast_launches.csv > $OUT/${TAG}_f16fast_launches.summary.txt 2>&1
  head -24 $OUT/${TAG}_f16fast_launches.summary.txt
  for K in edge_tc_kernel node_tc_kernel; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 3 \
        -o $OUT/${TAG}_f16fast_$K -f \
        python bench.py --steps 1 --warmup 1 --timesteps 8 --no-cpu-baseline > $OUT/${TAG}_ncu_$K.log 2>&1
    ncu -i $OUT/${TAG}_f16fast_$K.ncu-rep --page raw --csv > $OUT/${TAG}_f16fast_$K.raw.csv 2>/dev/null
    tail -2 $OUT/${TAG}_ncu_$K.log
  done
  for W in config3 config5; do
    timeout 600 ncu --set full --clock-control none -k regex:edge_tc_kernel -s 12 -c 2 \
        -o $OUT/${TAG}_${W}_edge_tc_kernel -f \
        python bench.py --workload $W --steps 1 --warmup 1 --timesteps 4 --no-cpu-baseline > $OUT/${TAG}_ncu_${W}.log 2>&1
    ncu -i $OUT/${TAG}_${W}_edge_tc_kernel.ncu-rep --page raw --csv > $OUT/${TAG}_${W}_edge_tc_kernel.raw.csv 2>/dev/null
  done
fi
ls -la $OUT | tail -40
