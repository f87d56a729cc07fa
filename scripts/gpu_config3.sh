#!/bin/bash
# config-3 / config-5 launch lists (ncu, cold-cache serialised: compare SHARES) + bench lines
set -u
TAG=${1:-r05n}
OUT=gpurun_out
mkdir -p $OUT
for W in config3 config5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv \
      --log-file $OUT/${TAG}_${W}_launches.csv \
      python bench.py --workload $W --steps 1 --warmup 1 --timesteps 6 --no-cpu-baseline > $OUT/${TAG}_${W}_ncu_launches.log 2>&1
  python profiles/summarize_launches.py $OUT/${TAG}_${W}_launches.csv > $OUT/${TAG}_${W}_launches.summary.txt 2>&1
  cat $OUT/${TAG}_${W}_launches.summary.txt
  timeout 900 python bench.py --workload $W --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench_${W}.json 2> $OUT/${TAG}_bench_${W}.err
  python -c "
import json
j = json.load(open('$OUT/${TAG}_bench_${W}.json')); print('$W', 'samples/s %.2f' % j['value'], 'step_us %.1f' % j['denoise_step_us'], 'msg frac %.3f' % j['roofline']['frac'], 'msg us %.1f' % j['roofline']['avg_launch_us'], {k: round(v, 2) for k, v in j['roofline']['kernel_ms_by_kind'].items()})"
done
