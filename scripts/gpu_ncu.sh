#!/bin/bash
# ncu --set full capture of one kernel (regex) from a short bench run.
#   gpurun --timeout 600 -- 'bash scripts/gpu_ncu.sh <tag> <kernel regex> [precision]'
TAG=${1:-n}; K=${2:-edge_tc_kernel}; PREC=${3:-bf16}
OUT=gpurun_out; mkdir -p $OUT
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 2 -o $OUT/${TAG}_${PREC}_$K -f \
    python bench.py --precision $PREC --steps 1 --warmup 1 --timesteps 8 --no-cpu-baseline > $OUT/${TAG}_ncu_$K.log 2>&1
tail -2 $OUT/${TAG}_ncu_$K.log
