#!/bin/bash
# round 2 (second session), call 17: per-kind register budgets at BASELINE's 2^18 envs (one round of tiles instead of two)
mkdir -p gpurun_out
for e in Acrobot-v1 MountainCar-v0 MountainCarContinuous-v0 Pendulum-v1; do
  timeout 300 python bench.py --env $e --log2-envs 18 --steps 400 --warmup 50 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('default $e ms', d['ms_per_step'], 'value %.3g' % d['value'])"
done
for e in Pendulum-v1 CartPole-v1; do
  B200GYM_LIB=$PWD/_variants/lib_c8.so timeout 300 python bench.py --env $e --log2-envs 18 --steps 400 --warmup 50 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('32-register build $e 2^18 ms', d['ms_per_step'], 'value %.3g' % d['value'])"
done
timeout 300 python bench.py --env CartPole-v1 --log2-envs 18 --steps 400 --warmup 50 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('default CartPole-v1 2^18 ms', d['ms_per_step'], 'value %.3g' % d['value'])"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -2
