#!/bin/bash
# round 2 (second session): final validation of the final tree on one GPU + a source-level ncu capture of the headline kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 > gpurun_out/r2b_final2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -9 gpurun_out/r2b_final2_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/r2b_final2_bench_n1.json 2> gpurun_out/r2b_final2_bench_n1.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_final2_bench_n1.json'))
print('n1 ms', d['ms_per_step'], 'value %.4g' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.3g' % d['e2e']['value'], 'launches', d['gpu_launches'], d['clocks'])
for k,v in d.get('configs', {}).items(): print(' ', k, v.get('ms_per_step'), '%.3g' % v.get('value', 0), v.get('error'))
PY
timeout 300 ncu --set full --import-source on --clock-control none -k regex:step_kernel_persistent -s 40 -c 1 -f -o gpurun_out/r2b_cartpole_kernel_final python bench.py --steps 20 --warmup 40 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2b_ncu_final.log 2>&1; echo "ncu rc=$?"
