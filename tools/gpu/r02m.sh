#!/bin/bash
# round 2, call m (1 GPU): RK operands staged by TMA (OPTMA) against per-thread cp.async
mkdir -p gpurun_out
timeout 300 python tools/sweep_march.py C4 0,11,12 3 > gpurun_out/r02m_sweep_c4_optma.txt 2>&1; cat gpurun_out/r02m_sweep_c4_optma.txt
timeout 300 python tools/sweep_march.py C4r3 0,11 3 >> gpurun_out/r02m_sweep_c4_optma.txt 2>&1; tail -2 gpurun_out/r02m_sweep_c4_optma.txt
HB_MARCH_CFG=11 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "C4_ or march3d or slab_march3d or slab_thin3d" > gpurun_out/r02m_pytest_optma.log 2>&1; tail -3 gpurun_out/r02m_pytest_optma.log
HB_MARCH_CFG=11 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_bench_c4_optma.json 2> gpurun_out/r02m_bench_c4_optma.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02m_bench_c4_optma.json').read().strip().splitlines()[-1]); print('C4 optma', '%.3f G/s'%(d['value']/1e9), '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'parity', (d.get('parity') or {}).get('rel_linf'), d['roofline']['kernel_config'][-80:])
except Exception as e: print('FAILED', e, open('gpurun_out/r02m_bench_c4_optma.err').read()[-500:])
PY
