#!/bin/bash
# round 2 (second session): final validation on one GPU -- what the driver runs at round end, plus the launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 > gpurun_out/r2b_final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -10 gpurun_out/r2b_final_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/r2b_final_bench_n1.json 2> gpurun_out/r2b_final_bench_n1.err ) 2>&1 | grep real; echo "bench rc=$?"
( time timeout 900 python bench.py --impl reference > gpurun_out/r2b_final_bench_reference.json 2> gpurun_out/r2b_final_bench_reference.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_final_bench_n1.json'))
print('n1 ms', d['ms_per_step'], 'value %.4g' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'e2e %.3g' % d['e2e']['value'], 'launches', d['gpu_launches'], d['clocks'])
print('cpu', d['cpu_baseline']['value'], 'py', json.dumps(d['cpu_baseline_python'])[:400])
for k,v in d.get('configs', {}).items(): print(' ', k, v.get('ms_per_step'), '%.3g' % v.get('value', 0), v.get('error'))
print('strong', d.get('strong_scaling', {}).get('ms_per_step'))
r=json.load(open('gpurun_out/r2b_final_bench_reference.json'))
print('reference arm', r.get('value'), r.get('cpu_baseline'))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1; echo "ncu launch list rc=$?"
python scripts/launch_list_summary.py gpurun_out/r2b_launches.csv | head -14
