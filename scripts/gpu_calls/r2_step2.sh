#!/bin/bash
# round 2, call 3 (2 GPUs): new host path + live-reference tests on GPU 0, full 2-GPU bench line, ncu of kernel P
NG=${1:-2}
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_multi.py --durations=6 > gpurun_out/r2_pytest_gpu_a.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest_gpu_a.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 300 --warmup 30 --cpu-seconds 3 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1_b.json'))
print('n1', d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e'], 'py', d['cpu_baseline_python'])
PY
for g in direct bulk; do
CUDA_VISIBLE_DEVICES=0 B200GYM_GATHER=$g timeout 200 python - <<'PY'
import time, numpy as np, torch, gym_b200, os
n=1<<20
env=gym_b200.vector.make("CartPole-v1", n, backend="numpy", copy=False, dense_infos=True); env.reset(seed=0)
acts=[torch.randint(0,2,(n,),dtype=torch.int64).pin_memory().numpy() for _ in range(4)]
for k in range(5): env.step(acts[k%4])
for ch in (1,2,4,8):
    os.environ["B200GYM_HOST_CHUNKS"]=str(ch)
    for k in range(3): env.step(acts[k%4])
    t0=time.perf_counter()
    for k in range(100): env.step(acts[k%4])
    el=time.perf_counter()-t0
    print(os.environ.get("B200GYM_GATHER"), "chunks", ch, "ms/step", 1e3*el/100, "env-steps/s %.3g" % (n*100/el))
PY
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 200 --warmup 30 2>/dev/null | grep '^{' > gpurun_out/r2_bench_n${NG}_b.json; echo "bench n$NG rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_n${NG}_b.json'))
print('n$NG', d['ms_per_step'], d['value'], d['gather_verified'], 'e2e', d['e2e'], 'strong', d['strong_scaling']['ms_per_step'])
PY
CUDA_VISIBLE_DEVICES=0 B200GYM_KERNEL=p timeout 300 ncu --set full --import-source on --clock-control none -k regex:step_kernel_persistent -s 40 -c 1 -f -o gpurun_out/r2_cartpole_kernel_P python bench.py --steps 20 --warmup 40 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2_ncu_P.log 2>&1; echo "ncu P rc=$?"
