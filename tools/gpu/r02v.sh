#!/bin/bash
# round 2, call v (1 GPU): planes per CTA (KM = 128 / 32 against 64) on the full 512^3 workload
mkdir -p gpurun_out
for c in 0 4 5; do
  HB_MARCH_CFG=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r02v_bench_cfg$c.json 2> gpurun_out/r02v_bench_cfg$c.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02v_bench_cfg$c.json').read().strip().splitlines()[-1]); print('cfg $c', '%.3f G/s'%(d['value']/1e9), '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], d['roofline']['kernel_config'][-100:])
except Exception as e: print('cfg $c FAILED', e, open('gpurun_out/r02v_bench_cfg$c.err').read()[-300:])
PY
done
