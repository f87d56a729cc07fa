#!/bin/bash
# compute-sanitizer over the small sampler workload (race / memory checking of the hand-rolled
# mbarrier rings, named barriers, setmaxnreg and TMEM allocation).  One gpurun call:
#   gpurun --timeout 1800 -- 'bash scripts/gpu_sanitize.sh r05s'
# memcheck and synccheck run every precision; racecheck runs with a raised hazard limit so that every hazard CLASS shows
# (racecheck does not model tcgen05.commit -> mbarrier ordering: see profiles/r05_sanitizer_notes.md).
set -u
TAG=${1:-r05s}
OUT=gpurun_out
mkdir -p $OUT
export NV_COMPUTE_SANITIZER_MAX_RACECHECK_HAZARDS=4000
for TOOL in ${SAN_TOOLS:-memcheck synccheck racecheck}; do
  for PREC in ${SAN_PRECS:-fp32 tf32 f16 f16fast bf16}; do
    timeout ${SAN_TIMEOUT:-300} compute-sanitizer --tool $TOOL --print-limit 4000 python scripts/sanitize_workload.py $PREC \
        > $OUT/${TAG}_sanitize_${TOOL}_${PREC}.txt 2>&1
    echo "$TOOL $PREC rc=$?" | tee -a $OUT/${TAG}_sanitize_summary.txt
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/${TAG}_sanitize_${TOOL}_${PREC}.txt | tee -a $OUT/${TAG}_sanitize_summary.txt
    if [ "$TOOL" = "racecheck" ]; then
      # hazard classes: (kernel function, source line) pairs with counts
      grep -E "Race reported|     and " $OUT/${TAG}_sanitize_${TOOL}_${PREC}.txt | sed -E 's/\+0x[0-9a-f]+//; s/void <unnamed>:://; s/\[[0-9]+ hazards\]//; s/\(.*\) in / in /' \
          | sort | uniq -c | sort -rn | head -40 >> $OUT/${TAG}_sanitize_summary.txt
      # keep the raw log small
      head -c 300000 $OUT/${TAG}_sanitize_${TOOL}_${PREC}.txt > $OUT/${TAG}_sanitize_${TOOL}_${PREC}.head.txt && rm $OUT/${TAG}_sanitize_${TOOL}_${PREC}.txt
    fi
  done
done
