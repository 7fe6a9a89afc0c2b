mkdir -p gpurun_out
for v in c8_s2 c8_s3 c6_s3 c5_s3; do
  for simple in 0 1; do
    B200GYM_LIB=$PWD/gym_b200/variants/lib_$v.so B200GYM_SIMPLE_KERNEL=$simple python bench.py --steps 300 --warmup 30 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v simple=$simple', round(d['ms_per_step']*1e3,2), 'us cold;', round(d['warm_l2']['ms_per_step']*1e3,2), 'us warm; frac', round(d['roofline']['frac'],3))"
  done
done
