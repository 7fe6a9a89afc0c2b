#!/bin/bash
# round 2: kernel P vs A on one GPU, kernel G (bulk pushes) vs per-thread peer stores on 2+ GPUs
NG=${1:-2}
mkdir -p gpurun_out
export CUDA_VISIBLE_DEVICES_ALL=$(seq -s, 0 $((NG-1)))
CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "persistent or free_run" 2>&1 | tail -3
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
for k in a p; do
  CUDA_VISIBLE_DEVICES=0 B200GYM_KERNEL=$k timeout 200 python bench.py --steps 400 --warmup 50 --no-cpu-baseline --no-e2e > gpurun_out/r2_ab_kernel_$k.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/r2_ab_kernel_$k.json'));print('kernel $k', d['ms_per_step'], d['roofline']['frac'], d.get('warm_l2'))"
done
for g in direct bulk; do
  B200GYM_GATHER=$g timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 300 --warmup 30 --no-e2e > gpurun_out/r2_ab_gather_${g}_n$NG.json 2> gpurun_out/r2_ab_gather_${g}_n$NG.err
  python -c "import json;d=json.load(open('gpurun_out/r2_ab_gather_${g}_n$NG.json'));print('gather $g n=$NG', d['ms_per_step'], d['value'])" || tail -5 gpurun_out/r2_ab_gather_${g}_n$NG.err
done
NCCL_DEBUG=INFO timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $NG --steps 50 --warmup 5 --no-e2e --gather nccl 2>&1 | grep -i -E "nvls|multicast|json|metric" | head -8 | cut -c1-300
nvidia-smi topo -m | head -12
