#!/bin/bash
# round 2 (second session), call 7: TOI kernels with few envs per warp (lunar + walker): parity, cost, ncu launch durations
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lunar.py tests/test_gpu_walker.py -q -m gpu > gpurun_out/r2b_pytest_gpu_toi5.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2b_pytest_gpu_toi5.log
B200GYM_BOX2D_TOI_DEFER=0 timeout 600 python -m pytest tests/test_gpu_walker.py -q -m gpu > gpurun_out/r2b_pytest_gpu_toi5_inline.log 2>&1; echo "pytest inline rc=$?"; tail -2 gpurun_out/r2b_pytest_gpu_toi5_inline.log
for e in LunarLander-v2 LunarLanderContinuous-v2 BipedalWalker-v3 BipedalWalkerHardcore-v3; do
  timeout 300 python bench.py --env $e --log2-envs 16 --steps 100 --warmup 250 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep '^{' > gpurun_out/r2b_box2d5_${e}.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r2b_box2d5_${e}.json'))
print('$e', 'ms', d['ms_per_step'], 'value %.3g' % d['value'])
PY
done
B200GYM_BOX2D_TOI_DEFER=0 timeout 300 python bench.py --env BipedalWalker-v3 --log2-envs 16 --steps 100 --warmup 250 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('BipedalWalker-v3 inline TOI ms', d['ms_per_step'])"
for e in LunarLander-v2 BipedalWalker-v3; do
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,launch__grid_size --clock-control none -k regex:'lunar_|walker_' -s 520 -c 4 --csv --log-file gpurun_out/r2b_launches_$e.csv python bench.py --env $e --log2-envs 16 --steps 20 --warmup 250 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1; echo "ncu $e rc=$?"
cut -d, -f5,12- gpurun_out/r2b_launches_$e.csv | tail -18
done
