#!/bin/bash
# round 2 (second session), call 5: bit-exact early exit of the TOI velocity iterations: Box2D parity + cost
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lunar.py tests/test_gpu_walker.py tests/test_gpu_api.py -q -m gpu > gpurun_out/r2b_pytest_gpu_toi3.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2b_pytest_gpu_toi3.log
for e in LunarLander-v2 BipedalWalker-v3 BipedalWalkerHardcore-v3 LunarLanderContinuous-v2; do
  timeout 300 python bench.py --env $e --log2-envs 16 --steps 100 --warmup 250 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep '^{' > gpurun_out/r2b_box2d3_$e.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r2b_box2d3_$e.json'))
print('$e', 'ms', d['ms_per_step'], 'value %.3g' % d['value'])
PY
done
timeout 400 ncu --set full --import-source on --clock-control none -k regex:lunar_step_kernel -s 260 -c 1 -f -o gpurun_out/r2b_lunar_toi3 python bench.py --env LunarLander-v2 --log2-envs 16 --steps 20 --warmup 250 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2b_ncu_lunar3.log 2>&1; echo "ncu lunar rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:walker_step_kernel -s 260 -c 1 -f -o gpurun_out/r2b_walker_toi3 python bench.py --env BipedalWalker-v3 --log2-envs 16 --steps 20 --warmup 250 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2b_ncu_walker3.log 2>&1; echo "ncu walker rc=$?"
