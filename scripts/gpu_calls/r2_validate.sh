#!/bin/bash
# round 2: full GPU test suite (multi-GPU tests included when >= 2 GPUs are visible) + a short N=1 bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2_pytest_gpu_full.log 2>&1; echo "pytest rc=$?"; tail -18 gpurun_out/r2_pytest_gpu_full.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --steps 300 --warmup 30 --cpu-seconds 3 > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1_c.json'))
print('n1 ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e %.3g' % d['e2e']['value'])
for k,v in d['configs'].items(): print(' ', k, v.get('ms_per_step'), '%.3g' % v.get('value', 0), v.get('error'))
PY
