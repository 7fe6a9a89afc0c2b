#!/bin/bash
# round 2, first GPU call: (1) can box2d-py be had on the GPU image?  (2) BASELINE config 3 bench lines + ncu captures
mkdir -p gpurun_out
{
  echo "== Box2D probe on the GPU image =="
  python -c "import Box2D; print('Box2D', Box2D.__version__)" 2>&1 | tail -1
  python -c "import box2d" 2>&1 | tail -1
  python -m pip download --no-deps -d /tmp/b2 box2d-py==2.3.5 2>&1 | tail -2
  python -m pip install swig box2d-py 2>&1 | tail -2
  ls /opt/wheelhouse 2>/dev/null | grep -i -E "box|swig|gym|pygame" ; echo "wheelhouse grep rc=$?"
  which swig; echo "swig rc=$?"
  find / -iname "*box2d*" -not -path "/proc/*" -not -path "$GRAFT_REPO_ROOT/*" 2>/dev/null | head; echo "find done"
  nproc; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv
} > gpurun_out/r2_box2d_probe.txt 2>&1
cat gpurun_out/r2_box2d_probe.txt
for e in Pendulum-v1 Acrobot-v1 MountainCar-v0 MountainCarContinuous-v0; do
  timeout 300 python bench.py --env $e --log2-envs 18 --steps 1000 --warmup 100 --cpu-seconds 5 > gpurun_out/r2_bench_${e}_2p18.json 2> gpurun_out/r2_bench_${e}.err; echo "$e rc=$?"
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:step_kernel -s 30 -c 1 -f -o gpurun_out/r2_${e}_step python bench.py --env $e --log2-envs 18 --steps 20 --warmup 40 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_${e}.log 2>&1; echo "ncu $e rc=$?"
done
timeout 300 python bench.py --steps 1000 --warmup 100 --cpu-seconds 5 > gpurun_out/r2_bench_cartpole_start.json 2>/dev/null; echo "cartpole rc=$?"
cat gpurun_out/r2_bench_*_2p18.json gpurun_out/r2_bench_cartpole_start.json | cut -c1-600
