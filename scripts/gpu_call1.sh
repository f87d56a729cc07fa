#!/bin/bash
# Round-2 validation call: every GPU test (no -x: all failures in one pass; one pytest process per file so a hang costs one
# file), smoke, a first bench line, compute-sanitizer over the small sampler workload.
set -u
TAG=${1:-r05a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
: > $OUT/${TAG}_pytest.txt
for F in tests/test_gpu_parity.py tests/test_large_golden.py tests/test_cli.py tests/test_analysis.py; do
  echo "=== $F" >> $OUT/${TAG}_pytest.txt
  timeout 600 python -m pytest $F -m gpu -q -s --timeout 240 --timeout-method thread >> $OUT/${TAG}_pytest.txt 2>&1
  echo "rc=$?" >> $OUT/${TAG}_pytest.txt
done
grep -E "passed|failed|FAILED|ERROR|Timeout|rc=|===" $OUT/${TAG}_pytest.txt | tail -60
grep "\[parity\]" $OUT/${TAG}_pytest.txt > $OUT/${TAG}_parity_errors.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1
tail -4 $OUT/${TAG}_smoke.txt
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
for TOOL in ${SAN_TOOLS:-memcheck synccheck racecheck}; do
  for PREC in ${SAN_PRECS:-f16fast}; do
    timeout ${SAN_TIMEOUT:-240} compute-sanitizer --tool $TOOL --print-limit 20 python scripts/sanitize_workload.py $PREC \
        > $OUT/${TAG}_sanitize_${TOOL}_${PREC}.txt 2>&1
    echo "$TOOL $PREC rc=$?" | tee -a $OUT/${TAG}_sanitize_summary.txt
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Error" $OUT/${TAG}_sanitize_${TOOL}_${PREC}.txt | head -8 | tee -a $OUT/${TAG}_sanitize_summary.txt
  done
done
