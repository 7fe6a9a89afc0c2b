#!/bin/bash
# round 2 (second session), call 2: register budgets / cold-path placement of the CartPole step kernel
mkdir -p gpurun_out
run() {  # label, lib, kernel, ctas
  B200GYM_LIB=$2 B200GYM_KERNEL=$3 B200GYM_P_CTAS=$4 timeout 200 python bench.py --steps 400 --warmup 50 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep '^{' > gpurun_out/r2b_v_$1.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r2b_v_$1.json'))
print('$1: median ms', d['ms_per_step'], 'mean', d.get('ms_per_step_mean'), 'p10', d.get('ms_per_step_p10'), 'frac', d['roofline']['frac'])
PY
}
D=$PWD/gym_b200/libb200gym.so
for k in a p l; do run default_$k $D $k -1; done
for k in a p l; do run cudasincos_$k $PWD/_variants/lib_cudasincos.so $k -1; done
for k in a p l; do run c6_$k $PWD/_variants/lib_c6.so $k -1; done
for k in a p l; do run c5_$k $PWD/_variants/lib_c5.so $k -1; done
