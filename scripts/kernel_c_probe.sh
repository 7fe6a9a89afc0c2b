#!/bin/bash
# kernel C (two envs per thread) vs kernel A: parity suite under C, then the CartPole 2^20 bench under both
B200GYM_KERNEL=c timeout 200 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
for k in a c a c; do
  B200GYM_KERNEL=$k timeout 100 python bench.py --steps 400 --warmup 50 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('kernel $k', 'ms/step', round(d['ms_per_step'],5), 'frac', round(d['roofline']['frac'],3), 'warm', round(d['warm_l2']['ms_per_step'],5))"
done
