#!/bin/bash
# round 2 (second session), call 18 (2 GPUs): NVLink tx / rx bytes per launch of the fused step kernel (kernel G)
mkdir -p gpurun_out
B200GYM_P2P_TIMEOUT_S=20 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 --no-python bash scripts/nvlink_counters.sh --gpus 2 --steps 30 --warmup 10 --no-e2e --no-extra > gpurun_out/r2b_nvlink.log 2>&1; echo "rc=$?"
grep -E "nvl|duration" gpurun_out/r2b_nvlink_rank0.csv | cut -d, -f5,12- | head -16
tail -3 gpurun_out/r2b_nvlink.log | cut -c1-300
