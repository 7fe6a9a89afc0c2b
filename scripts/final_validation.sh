#!/bin/bash
# round-end style validation on one B200: GPU tests, smoke, bench (+ reference arm), ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_multi.py --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r1.log 2> gpurun_out/bench_r1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 100 --warmup 5 > gpurun_out/bench_ref_r1.log 2>/dev/null; echo "ref rc=$?"
for e in LunarLander-v2 LunarLanderContinuous-v2 BipedalWalker-v3 BipedalWalkerHardcore-v3; do
  timeout 300 python bench.py --env $e --log2-envs 16 --steps 200 --warmup 20 --cpu-seconds 5 > gpurun_out/bench_$e.log 2>/dev/null; echo "$e rc=$?"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1; echo "list rc=$?"
