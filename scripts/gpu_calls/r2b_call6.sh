#!/bin/bash
# round 2 (second session), call 6: LunarLander with the parked-env TOI kernel vs inline TOI; parity; one ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lunar.py tests/test_gpu_walker.py -q -m gpu > gpurun_out/r2b_pytest_gpu_toi4.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2b_pytest_gpu_toi4.log
B200GYM_BOX2D_TOI_DEFER=0 timeout 600 python -m pytest tests/test_gpu_lunar.py -q -m gpu > gpurun_out/r2b_pytest_gpu_toi4_inline.log 2>&1; echo "pytest inline rc=$?"; tail -2 gpurun_out/r2b_pytest_gpu_toi4_inline.log
for d in 1 0; do for e in LunarLander-v2 LunarLanderContinuous-v2; do
  B200GYM_BOX2D_TOI_DEFER=$d timeout 300 python bench.py --env $e --log2-envs 16 --steps 100 --warmup 250 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep '^{' > gpurun_out/r2b_box2d4_${e}_$d.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r2b_box2d4_${e}_$d.json'))
print('$e', 'toi_defer=$d', 'ms', d['ms_per_step'], 'value %.3g' % d['value'])
PY
done; done
timeout 400 ncu --set full --clock-control none -k regex:lunar_ -s 520 -c 2 -f -o gpurun_out/r2b_lunar_toi4 python bench.py --env LunarLander-v2 --log2-envs 16 --steps 20 --warmup 250 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2b_ncu_lunar4.log 2>&1; echo "ncu lunar rc=$?"
ls -la gpurun_out/
