#!/bin/bash
# parity probe + device-resident timing of the four Box2D variants (one GPU)
PYTHONPATH=. timeout 300 python scripts/variant_probe.py 2>&1 | grep -E "ok=|first_bad|probe done|mismatch|Error|error" | tail -20
for e in LunarLander-v2 LunarLanderContinuous-v2 BipedalWalker-v3 BipedalWalkerHardcore-v3; do
  timeout 200 python bench.py --env $e --log2-envs 16 --steps 200 --warmup 20 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$e', 'ms/step', round(d['ms_per_step'],4), 'value %.3e' % d['value'])"
done
