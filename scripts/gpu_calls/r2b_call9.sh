#!/bin/bash
# round 2 (second session), call 9: one vs two tiles in flight per thread in the persistent step kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "persistent" 2>&1 | tail -2
for d in 1 2; do
  B200GYM_P_DEPTH=$d timeout 300 python bench.py --steps 400 --warmup 50 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' > gpurun_out/r2b_depth_$d.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r2b_depth_$d.json'))
print('depth $d: median ms', d['ms_per_step'], 'mean', d.get('ms_per_step_mean'), 'p10', d.get('ms_per_step_p10'), 'frac', d['roofline']['frac'], 'warm', d.get('warm_l2', {}).get('ms_per_step'), 'strong 2^23', d.get('strong_scaling', {}).get('ms_per_step'))
for k,v in d.get('configs', {}).items():
    if 'Lunar' not in k and 'Walker' not in k: print('   ', k, v.get('ms_per_step'))
PY
done
