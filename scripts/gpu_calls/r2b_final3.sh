#!/bin/bash
# round 2 (second session): last check of the final tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_final3_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b_final3_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 300 --warmup 30 --cpu-seconds 3 > gpurun_out/r2b_final3_bench_n1.json 2> gpurun_out/r2b_final3_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_final3_bench_n1.json'))
print('n1 ms', d['ms_per_step'], 'value %.4g' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.3g' % d['e2e']['value'])
for k,v in d.get('configs', {}).items(): print(' ', k, v.get('ms_per_step'), '%.3g' % v.get('value', 0), v.get('error'))
PY
