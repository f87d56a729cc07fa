#!/bin/bash
# Multi-GPU bench: bash scripts/gpu_multi.sh <tag> <n_gpus>   (under gpurun --gpus N)
set -u
TAG=${1:-r05m}; N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/${TAG}_smi.txt 2>&1
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 5 --warmup 3 > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err
tail -c 3000 $OUT/${TAG}_bench_${N}gpu.json; tail -5 $OUT/${TAG}_bench_${N}gpu.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("$OUT/${TAG}_bench_${N}gpu.json") if l.startswith("{")][-1])
    print("N=$N value", j["value"], "e2e", j["e2e"]["value"], "step_us", j["denoise_step_us"])
    for a in j.get("also", []):
        print(a)
except Exception as e:
    print("parse failed", e)
PY
