#!/bin/bash
# round 2 (second session), call 16: configs[2] kernels with the vectorised trig-table loads: times + ncu summaries (text only)
mkdir -p gpurun_out
for e in Acrobot-v1 Pendulum-v1 MountainCar-v0 MountainCarContinuous-v0; do
  timeout 300 python bench.py --env $e --log2-envs 18 --steps 400 --warmup 50 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep '^{' > gpurun_out/r2b_cfg2_$e.json
  python -c "import json; d=json.load(open('gpurun_out/r2b_cfg2_$e.json')); print('$e ms', d['ms_per_step'], 'value %.3g' % d['value'], 'warm', d.get('warm_l2', {}).get('ms_per_step'))"
done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -2
for e in Acrobot-v1 Pendulum-v1; do
timeout 300 ncu --set full --clock-control none -k regex:step_kernel -s 30 -c 1 -f -o /tmp/r2b_$e python bench.py --env $e --log2-envs 18 --steps 20 --warmup 30 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1; echo "ncu $e rc=$?"
python scripts/ncu_summary.py /tmp/r2b_$e.ncu-rep > gpurun_out/r2b_${e}_step_kernel_ncu_full.txt 2>&1
tail -24 gpurun_out/r2b_${e}_step_kernel_ncu_full.txt | grep -E "duration|registers|inst_executed.sum|issue_active|fp64|stalled|warps_active"
done
