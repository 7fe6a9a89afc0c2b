#!/bin/bash
# round 2 (second session), call 12: step-kernel resets on a side stream under the TOI kernel: parity + cost
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lunar.py tests/test_gpu_walker.py tests/test_gpu_wrappers.py -q -m gpu > gpurun_out/r2b_pytest_gpu_split.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b_pytest_gpu_split.log
for sp in 0 1; do for e in LunarLander-v2 BipedalWalker-v3 BipedalWalkerHardcore-v3 LunarLanderContinuous-v2; do
  B200GYM_BOX2D_SPLIT=$sp timeout 300 python bench.py --env $e --log2-envs 16 --steps 100 --warmup 250 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$e split=$sp ms', d['ms_per_step'], 'value %.3g' % d['value'])"
done; done
