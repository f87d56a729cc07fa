cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for P in 1 0; do DIFFPHAR_PDL=$P python bench.py --precision bf16 --no-cpu-baseline --steps 3 > gpurun_out/r01l_bench_pdl$P.json 2>gpurun_out/r01l_bench_pdl$P.err; python -c "
import json; d=json.load(open('gpurun_out/r01l_bench_pdl$P.json')); print('PDL=$P', d['value'], d['denoise_step_us'])"; tail -2 gpurun_out/r01l_bench_pdl$P.err; done
