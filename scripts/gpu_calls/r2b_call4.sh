#!/bin/bash
# round 2 (second session), call 4: warp-synchronous TOI rounds + tight reject: parity + cost; live-reference tests (oracle/_ref present)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/r2b_pytest_gpu_toi2.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2b_pytest_gpu_toi2.log
for e in LunarLander-v2 BipedalWalker-v3 BipedalWalkerHardcore-v3 LunarLanderContinuous-v2; do
  timeout 300 python bench.py --env $e --log2-envs 16 --steps 100 --warmup 250 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep '^{' > gpurun_out/r2b_box2d_$e.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r2b_box2d_$e.json'))
print('$e', 'ms', d['ms_per_step'], 'value %.3g' % d['value'])
PY
done
timeout 400 ncu --set full --import-source on --clock-control none -k regex:lunar_step_kernel -s 260 -c 1 -f -o gpurun_out/r2b_lunar_toi2 python bench.py --env LunarLander-v2 --log2-envs 16 --steps 20 --warmup 250 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2b_ncu_lunar2.log 2>&1; echo "ncu lunar rc=$?"
