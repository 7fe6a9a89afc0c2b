#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "libm or free_run" 2>&1 | tail -15
timeout 200 python scripts/acrobot_diag.py Acrobot-v1 2>&1 | tail -20
timeout 200 python scripts/acrobot_diag.py Pendulum-v1 2>&1 | tail -8
