#!/bin/bash
# round 2 (second session), call 10: Box2D step after the friction-drift exit: TOI kernel grid sweep + launch durations
mkdir -p gpurun_out
for e in LunarLander-v2 BipedalWalker-v3; do for tg in 4 12 32; do
  B200GYM_TOI_GRID=$tg timeout 300 python bench.py --env $e --log2-envs 16 --steps 100 --warmup 250 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$e toi_grid=$tg ms', d['ms_per_step'])"
done; done
for e in LunarLanderContinuous-v2 BipedalWalkerHardcore-v3; do
  timeout 300 python bench.py --env $e --log2-envs 16 --steps 100 --warmup 250 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$e ms', d['ms_per_step'])"
done
for e in LunarLander-v2 BipedalWalker-v3; do
timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:'lunar_|walker_' -s 780 -c 3 --csv --log-file gpurun_out/r2b_launches3_$e.csv python bench.py --env $e --log2-envs 16 --steps 20 --warmup 250 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1; echo "ncu $e rc=$?"
grep -E "duration|inst_executed" gpurun_out/r2b_launches3_$e.csv | cut -d, -f5,12-
done
