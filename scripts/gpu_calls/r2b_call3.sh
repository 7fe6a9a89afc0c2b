#!/bin/bash
# round 2 (second session), call 3: TOI on the device (Box2D parity tests + cost), new CartPole default (lean kernel, 40 registers)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2b_pytest_gpu_toi.log 2>&1; echo "pytest rc=$?"; tail -16 gpurun_out/r2b_pytest_gpu_toi.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 300 --warmup 30 --cpu-seconds 3 > gpurun_out/r2b_bench_n1_toi.json 2> gpurun_out/r2b_bench_n1_toi.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench_n1_toi.json'))
print('n1 ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e %.3g' % d['e2e']['value'])
for k,v in d.get('configs', {}).items(): print(' ', k, v.get('ms_per_step'), '%.3g' % v.get('value', 0), v.get('error'))
PY
timeout 300 ncu --set full --import-source on --clock-control none -k regex:step_kernel_persistent -s 40 -c 1 -f -o gpurun_out/r2b_cartpole_kernel_L40 python bench.py --steps 20 --warmup 40 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2b_ncu_L40.log 2>&1; echo "ncu L40 rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:lunar_step_kernel -s 260 -c 1 -f -o gpurun_out/r2b_lunar_toi python bench.py --env LunarLander-v2 --log2-envs 16 --steps 20 --warmup 250 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2b_ncu_lunar.log 2>&1; echo "ncu lunar rc=$?"
