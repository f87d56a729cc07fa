#!/bin/bash
# compute-sanitizer over the small sampler workload (race / memory checking of the hand-rolled
# mbarrier rings, named barriers, setmaxnreg and TMEM allocation).  One gpurun call:
#   gpurun --timeout 1500 -- 'bash scripts/gpu_sanitize.sh r05'
set -u
TAG=${1:-r05}
OUT=gpurun_out
mkdir -p $OUT
for TOOL in memcheck synccheck racecheck initcheck; do
  for PREC in ${SAN_PRECS:-fp32 f16fast bf16}; do
    timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $TOOL --print-limit 20 python scripts/sanitize_workload.py $PREC \
        > $OUT/${TAG}_sanitize_${TOOL}_${PREC}.txt 2>&1
    echo "$TOOL $PREC rc=$?" | tee -a $OUT/${TAG}_sanitize_summary.txt
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Error" $OUT/${TAG}_sanitize_${TOOL}_${PREC}.txt | head -8 | tee -a $OUT/${TAG}_sanitize_summary.txt
  done
done
