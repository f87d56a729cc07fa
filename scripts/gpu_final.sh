#!/bin/bash
# End-of-round validation in one gpurun call: build check, full GPU tests, smoke, the default bench line, ncu launch list,
# ncu --set full of the edge (message + coordinate launches) and node kernels at config 2 and of the message kernel at
# config 3 (raw pages exported to CSV on the box), compute-sanitizer on the small workload.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_final.sh <tag>'
set -u
TAG=${1:-r06z}; PREC=f16fast
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1; tail -3 $OUT/${TAG}_smoke.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.txt 2>&1; tail -3 $OUT/${TAG}_pytest.txt
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json | cut -c1-1500
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>> $OUT/${TAG}_bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500 --csv --log-file $OUT/${TAG}_${PREC}_launches.csv \
      python bench.py --precision $PREC --steps 1 --warmup 1 --timesteps 30 --no-cpu-baseline --no-also > $OUT/${TAG}_ncu_launches.log 2>&1
  python profiles/summarize_launches.py $OUT/${TAG}_${PREC}_launches.csv > $OUT/${TAG}_${PREC}_launches.summary.txt 2>&1
  cat $OUT/${TAG}_${PREC}_launches.summary.txt | head -14
  for K in edge_tc_kernel node_tc_kernel; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 3 -o $OUT/${TAG}_${PREC}_$K -f \
        python bench.py --precision $PREC --steps 1 --warmup 1 --timesteps 8 --no-cpu-baseline --no-also > $OUT/${TAG}_ncu_$K.log 2>&1
    ncu -i $OUT/${TAG}_${PREC}_$K.ncu-rep --page raw --csv > $OUT/${TAG}_${PREC}_$K.raw.csv 2>/dev/null
    python scripts/ncu_raw_summary.py $OUT/${TAG}_${PREC}_$K.raw.csv > $OUT/${TAG}_${PREC}_$K.ncu_raw.txt 2>&1
  done
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:edge_tc_kernel -s 12 -c 1 -o $OUT/${TAG}_config3_edge_tc_kernel -f \
      python bench.py --precision $PREC --workload config3 --steps 1 --warmup 1 --timesteps 4 --no-cpu-baseline --no-also > $OUT/${TAG}_ncu_config3.log 2>&1
  ncu -i $OUT/${TAG}_config3_edge_tc_kernel.ncu-rep --page raw --csv > $OUT/${TAG}_config3_edge_tc_kernel.raw.csv 2>/dev/null
  python scripts/ncu_raw_summary.py $OUT/${TAG}_config3_edge_tc_kernel.raw.csv > $OUT/${TAG}_config3_edge_tc_kernel.ncu_raw.txt 2>&1
  grep -E "## launch|gpu__time_duration|dram__bytes|hmma|inst_issued" $OUT/${TAG}_*ncu_raw.txt | head -40
  rm -f $OUT/${TAG}_*.ncu-rep            # the reports are 10s of MB each: the raw CSV pages are what gets read
fi
if [ "${SKIP_SAN:-0}" != "1" ]; then
  SAN_PRECS="${SAN_PRECS:-f16fast bf16}" SAN_TIMEOUT=240 bash scripts/gpu_sanitize.sh $TAG
fi
