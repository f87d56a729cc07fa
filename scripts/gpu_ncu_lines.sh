#!/bin/bash
# ncu --set full captures of the message, coordinate and node kernels at config 2 and of the message kernel at config 3,
# summarised per source line ON THE BOX (scripts/ncu_lines.py needs the same build of the library):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_ncu_lines.sh <tag>'
TAG=${1:-r07n}; PREC=f16fast
OUT=gpurun_out; mkdir -p $OUT
cap() {  # name kernel-regex skip workload timesteps
  timeout 500 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c 1 -o $OUT/${TAG}_$1 -f \
      python bench.py --precision $PREC --workload $4 --steps 1 --warmup 1 --timesteps $5 --no-cpu-baseline --no-also > $OUT/${TAG}_ncu_$1.log 2>&1
  python scripts/ncu_lines.py $OUT/${TAG}_$1.ncu-rep "$6" 1 45 > $OUT/${TAG}_$1.lines.txt 2>&1
  head -12 $OUT/${TAG}_$1.lines.txt
  rm -f $OUT/${TAG}_$1.ncu-rep
}
cap config2_edge_msg   "edge_tc_kernel" 30 config2 8 edge_tc_kernel
cap config2_node       "node_tc_kernel"      13 config2 8 node_tc_kernel
cap config3_edge_msg   "edge_tc_kernel" 10 config3 4 edge_tc_kernel
