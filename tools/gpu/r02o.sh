#!/bin/bash
# round 2, call o (1 GPU): the folded last stage (running sum carried from stage to stage) against the unfolded plan
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ops.py -m gpu -q -x -k "rk_fold or C4_ or march or ops" > gpurun_out/r02o_pytest_fold.log 2>&1; tail -5 gpurun_out/r02o_pytest_fold.log
HB_RK_FOLD=0 timeout 300 python tools/sweep_march.py C4 0 3 > gpurun_out/r02o_sweep_fold.txt 2>&1
timeout 300 python tools/sweep_march.py C4 0 3 >> gpurun_out/r02o_sweep_fold.txt 2>&1
timeout 300 python tools/sweep_march.py M3r4 0 3 >> gpurun_out/r02o_sweep_fold.txt 2>&1
HB_RK_FOLD=0 timeout 300 python tools/sweep_march.py M3r4 0 3 >> gpurun_out/r02o_sweep_fold.txt 2>&1
cat gpurun_out/r02o_sweep_fold.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02o_bench_c4_fold.json 2> gpurun_out/r02o_bench_c4_fold.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02o_bench_c4_fold.json').read().strip().splitlines()[-1]); print('C4 fold', '%.3f G/s'%(d['value']/1e9), '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'parity', (d.get('parity') or {}).get('rel_linf'), d['roofline']['kernel_config'][-90:])
except Exception as e: print('FAILED', e, open('gpurun_out/r02o_bench_c4_fold.err').read()[-500:])
PY
