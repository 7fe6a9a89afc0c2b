#!/bin/bash
# round 2 (second session), call 15: LunarLander CTA size with the three-launch step; ncu of the Acrobot / Pendulum step kernels (glibc-exact trig)
mkdir -p gpurun_out
for bs in 128 256; do
  B200GYM_BOX2D_BLOCK=$bs timeout 300 python bench.py --env LunarLander-v2 --log2-envs 16 --steps 100 --warmup 250 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('LunarLander-v2 block=$bs ms', d['ms_per_step'])"
done
for e in Acrobot-v1 Pendulum-v1; do
timeout 300 ncu --set full --import-source on --clock-control none -k regex:step_kernel -s 30 -c 1 -f -o gpurun_out/r2b_$e python bench.py --env $e --log2-envs 18 --steps 20 --warmup 30 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2b_ncu_$e.log 2>&1; echo "ncu $e rc=$?"
done
ls -la gpurun_out/*.ncu-rep
