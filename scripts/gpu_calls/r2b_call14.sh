#!/bin/bash
# round 2 (second session), call 14: CartPole step with one validity flag instead of per-operation branches: parity + time
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -q -m gpu 2>&1 | tail -3
for k in l p a; do
  B200GYM_KERNEL=$k timeout 300 python bench.py --steps 400 --warmup 50 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('kernel $k: median ms', d['ms_per_step'], 'mean', d.get('ms_per_step_mean'), 'p10', d.get('ms_per_step_p10'), 'frac', d['roofline']['frac'], 'warm', d.get('warm_l2', {}).get('ms_per_step'))"
done
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:step_kernel_persistent -s 40 -c 2 python bench.py --steps 20 --warmup 40 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep -E "duration|inst_executed"
