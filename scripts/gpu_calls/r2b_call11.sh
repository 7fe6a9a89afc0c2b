#!/bin/bash
# round 2 (second session), call 11 (2 GPUs): multi-GPU tests + the 2-GPU bench line after this session's kernel changes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2b_pytest_multi_n2.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/r2b_pytest_multi_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 300 --warmup 50 2>gpurun_out/r2b_bench_n2.err | grep '^{' > gpurun_out/r2b_bench_n2.json; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench_n2.json'))
print('n2 ms', d['ms_per_step'], 'value %.4g' % d['value'], 'verified', d.get('gather_verified'), 'e2e', d['e2e'] and '%.3g' % d['e2e']['value'], 'strong', d.get('strong_scaling', {}).get('ms_per_step'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('reference arm n2', d['value'], d['config']['workload'])"
