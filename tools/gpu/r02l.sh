#!/bin/bash
# round 2, call l (1 GPU): RK operands read directly in the epilogue (OPDIRECT) against cp.async staging
mkdir -p gpurun_out
timeout 300 python tools/sweep_march.py C4 0,10,11 3 > gpurun_out/r02l_sweep_c4_opdirect.txt 2>&1; cat gpurun_out/r02l_sweep_c4_opdirect.txt
timeout 300 python tools/sweep_march.py C4r3 0,10 3 >> gpurun_out/r02l_sweep_c4_opdirect.txt 2>&1; tail -2 gpurun_out/r02l_sweep_c4_opdirect.txt
HB_MARCH_CFG=10 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "C4_ or march3d or slab_march3d or slab_thin3d" > gpurun_out/r02l_pytest_opdirect.log 2>&1; tail -3 gpurun_out/r02l_pytest_opdirect.log
HB_MARCH_CFG=10 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_bench_c4_opdirect.json 2> gpurun_out/r02l_bench_c4_opdirect.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02l_bench_c4_opdirect.json').read().strip().splitlines()[-1]); print('C4 opdirect', '%.3f G/s'%(d['value']/1e9), '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'parity', (d.get('parity') or {}).get('rel_linf'), d['roofline']['kernel_config'][-80:])
PY
