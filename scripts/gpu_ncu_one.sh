#!/bin/bash
# ncu --set full of one kernel: bash scripts/gpu_ncu_one.sh <tag> <kernel regex> <workload> [skip] [count]
set -u
TAG=$1; K=$2; W=${3:-config3}; S=${4:-2}; C=${5:-2}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -o $OUT/${TAG}_${W}_$K -f \
    python bench.py --workload $W --steps 1 --warmup 1 --timesteps 3 --no-cpu-baseline > $OUT/${TAG}_ncu_$K.log 2>&1
tail -3 $OUT/${TAG}_ncu_$K.log
ncu -i $OUT/${TAG}_${W}_$K.ncu-rep --page raw --csv > $OUT/${TAG}_${W}_$K.raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_${W}_$K.ncu-rep --page source --csv > $OUT/${TAG}_${W}_$K.source.csv 2>/dev/null
ls -la $OUT/${TAG}_${W}_$K.*
