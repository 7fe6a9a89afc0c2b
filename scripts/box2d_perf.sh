#!/bin/bash
# parity probe + device-resident timing of the Box2D variants (one GPU); deferred-reset A/B
PYTHONPATH=. timeout 300 python scripts/variant_probe.py 2>&1 | grep -E "ok=|first_bad|probe done|mismatch|Error|error" | tail -20
run() { timeout 200 python bench.py --env $1 --log2-envs ${2:-16} --steps 200 --warmup 20 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 2^${2:-16} defer=$B200GYM_BOX2D_DEFER', 'ms/step', round(d['ms_per_step'],4), 'value %.3e' % d['value'])"; }
for df in 0 1; do
  export B200GYM_BOX2D_DEFER=$df
  run LunarLander-v2
  run BipedalWalker-v3
done
unset B200GYM_BOX2D_DEFER
run LunarLanderContinuous-v2
run BipedalWalkerHardcore-v3
run LunarLander-v2 18
