#!/bin/bash
# round 2 (second session), call 13 (2 GPUs): the exchange's step barrier fused into kernel G's tail: parity + 2-GPU bench, fused vs separate
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2b_pytest_multi_fused.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/r2b_pytest_multi_fused.log
for f in 1 0; do
B200GYM_P2P_FUSE_BARRIER=$f timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$f bench.py --gpus 2 --steps 300 --warmup 50 --no-e2e 2>gpurun_out/r2b_bench_n2_f$f.err | grep '^{' > gpurun_out/r2b_bench_n2_f$f.json; echo "bench n2 fuse=$f rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2b_bench_n2_f$f.json'))
print('fuse=$f n2 ms', d['ms_per_step'], 'median', d.get('ms_per_step_median'), 'value %.4g' % d['value'], 'verified', d.get('gather_verified'), 'launches', d['gpu_launches'], 'strong', d.get('strong_scaling', {}).get('ms_per_step'))
PY
done
