#!/bin/bash
# round 2 (second session), call 1: validate the restored tree, then A/B the lean / balanced-grid variants of kernel P
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -16 gpurun_out/r2b_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
for k in p l; do for c in -1 0 7; do
  B200GYM_KERNEL=$k B200GYM_P_CTAS=$c timeout 200 python bench.py --steps 400 --warmup 50 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep '^{' > gpurun_out/r2b_ab_${k}_${c}.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r2b_ab_${k}_${c}.json'))
print('kernel $k ctas $c: median ms', d['ms_per_step'], 'mean', d.get('ms_per_step_mean'), 'p10', d.get('ms_per_step_p10'), 'frac', d['roofline']['frac'])
PY
done; done
B200GYM_KERNEL=l B200GYM_P_CTAS=0 timeout 300 ncu --set full --import-source on --clock-control none -k regex:step_kernel_persistent -s 40 -c 1 -f -o gpurun_out/r2b_cartpole_kernel_L python bench.py --steps 20 --warmup 40 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2b_ncu_L.log 2>&1; echo "ncu L rc=$?"
timeout 500 python bench.py --steps 300 --warmup 30 --cpu-seconds 3 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench_n1.json'))
print('n1 ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e %.3g' % d['e2e']['value'])
for k,v in d.get('configs', {}).items(): print(' ', k, v.get('ms_per_step'), '%.3g' % v.get('value', 0), v.get('error'))
PY
