#!/bin/bash
# One gpurun call: GPU parity tests, bench lines, ncu launch list and ncu --set full captures.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag> [precision]'
# Outputs land in gpurun_out/<tag>_*; copy the summaries you want judged into profiles/.
set -u
TAG=${1:-r01}
PREC=${2:-f16fast}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.txt 2>&1

if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1
  tail -5 $OUT/${TAG}_pytest.txt
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1
tail -2 $OUT/${TAG}_smoke.txt

timeout 600 python bench.py --precision $PREC > $OUT/${TAG}_bench_${PREC}.json 2> $OUT/${TAG}_bench_${PREC}.err
cat $OUT/${TAG}_bench_${PREC}.json
for P in ${EXTRA_PREC:-}; do
  timeout 600 python bench.py --precision $P --no-cpu-baseline > $OUT/${TAG}_bench_${P}.json 2> $OUT/${TAG}_bench_${P}.err
  cat $OUT/${TAG}_bench_${P}.json
done

if [ "${SKIP_NCU:-0}" != "1" ]; then
  # every launch with its device time (cold-cache, serialised: compare SHARES)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv \
      --log-file $OUT/${TAG}_${PREC}_launches.csv \
      python bench.py --precision $PREC --steps 1 --warmup 1 --timesteps 20 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
  python profiles/summarize_launches.py $OUT/${TAG}_${PREC}_launches.csv > $OUT/${TAG}_${PREC}_launches.summary.txt 2>&1
  cat $OUT/${TAG}_${PREC}_launches.summary.txt | head -20
  # the top kernels, full set
  for K in ${NCU_KERNELS:-edge_tc_kernel node_tc_kernel}; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 3 \
        -o $OUT/${TAG}_${PREC}_$K -f \
        python bench.py --precision $PREC --steps 1 --warmup 1 --timesteps 8 --no-cpu-baseline > $OUT/${TAG}_ncu_$K.log 2>&1
    tail -2 $OUT/${TAG}_ncu_$K.log
  done
fi
ls -la $OUT | tail -30
