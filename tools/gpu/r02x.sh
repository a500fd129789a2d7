#!/bin/bash
# round 2, call x (1 GPU): the shipped binary -- whole GPU suite, smoke, default bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02x_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02x_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02x_smoke.txt 2>&1; tail -1 gpurun_out/r02x_smoke.txt
timeout 600 python bench.py > gpurun_out/r02x_bench_default.json 2> gpurun_out/r02x_bench_default.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02x_bench_default.json').read().strip().splitlines()[-1]); print('default', '%.3f G/s'%(d['value']/1e9), '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.3f (blocking %.3f)'%(d['e2e']['value']/1e9, d['e2e']['blocking']['value']/1e9), 'launches', d['gpu_launches'], 'parity', d['parity']['rel_linf'], d['parity']['pass'])
except Exception as e: print('FAILED', e)
PY
