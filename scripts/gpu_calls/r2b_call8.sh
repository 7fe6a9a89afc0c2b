#!/bin/bash
# round 2 (second session), call 8: TOI kernel grid / deferred resets sweep (LunarLander, BipedalWalker)
mkdir -p gpurun_out
for e in LunarLander-v2 BipedalWalker-v3; do for rd in 0 1; do for tg in 12 64 256; do
  B200GYM_BOX2D_DEFER=$rd B200GYM_TOI_GRID=$tg timeout 300 python bench.py --env $e --log2-envs 16 --steps 100 --warmup 250 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$e reset_defer=$rd toi_grid=$tg ms', d['ms_per_step'])"
done; done; done
B200GYM_BOX2D_DEFER=1 timeout 600 python -m pytest tests/test_gpu_lunar.py tests/test_gpu_walker.py -q -m gpu -x 2>&1 | tail -2
for e in LunarLander-v2; do
B200GYM_BOX2D_DEFER=1 timeout 400 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,launch__grid_size --clock-control none -k regex:'lunar_|walker_' -s 780 -c 3 --csv --log-file gpurun_out/r2b_launches2_$e.csv python bench.py --env $e --log2-envs 16 --steps 20 --warmup 250 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1; echo "ncu $e rc=$?"
grep duration gpurun_out/r2b_launches2_$e.csv | cut -d, -f1,12-
done
