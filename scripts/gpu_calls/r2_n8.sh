#!/bin/bash
# round 2: the 8-GPU evidence run -- multi-GPU correctness at world=8, then the bench with both gather flavours
NG=${1:-8}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_pytest_multi_n$NG.log 2>&1; echo "pytest multi rc=$?"; tail -4 gpurun_out/r2_pytest_multi_n$NG.log
for g in bulk direct; do
  B200GYM_GATHER=$g timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 200 --warmup 30 $( [ $g = direct ] && echo --no-e2e ) 2>gpurun_out/r2_bench_n${NG}_$g.err | grep '^{' > gpurun_out/r2_bench_n${NG}_$g.json; echo "bench $g rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_n${NG}_$g.json'))
print('$g n$NG ms/step', d['ms_per_step'], 'value %.4g' % d['value'], 'verified', d['gather_verified'], 'median', d.get('ms_per_step_median'), 'strong', d['strong_scaling']['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('value'))
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $NG --steps 100 --warmup 10 --no-e2e --no-extra --gather nccl 2>/dev/null | grep '^{' > gpurun_out/r2_bench_n${NG}_nccl.json
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_n${NG}_nccl.json'));print('nccl n$NG ms/step', d['ms_per_step'], d['value'], d['gather_verified'])"
