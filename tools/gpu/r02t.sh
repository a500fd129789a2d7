#!/bin/bash
# round 2, call t (1 GPU): whole GPU suite with the 'plm eig' / 'plm eig prim' / 'plm eig prim ref' cases; default bench (e2e warm-up fix); tile-kernel timing check
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02t_pytest_gpu.log 2>&1
tail -8 gpurun_out/r02t_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02t_smoke.txt 2>&1; tail -1 gpurun_out/r02t_smoke.txt
timeout 600 python bench.py --workload C2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02t_bench_C2.json 2> gpurun_out/r02t_bench_C2.err
timeout 600 python bench.py > gpurun_out/r02t_bench_default.json 2> gpurun_out/r02t_bench_default.err
python - <<PY
import json
for w in ("C2","default"):
    try:
        d=json.loads(open('gpurun_out/r02t_bench_%s.json'%w).read().strip().splitlines()[-1]); print(w, '%.3f G/s'%(d['value']/1e9), '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.3f (blocking %.3f)'%(d['e2e']['value']/1e9, d['e2e']['blocking']['value']/1e9), 'launches', d['gpu_launches'])
    except Exception as e: print(w, 'FAILED', e)
PY
timeout 300 python bench.py --workload C4FL --steps 5 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r02t_bench_C4FL.json 2> gpurun_out/r02t_bench_C4FL.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02t_bench_C4FL.json').read().strip().splitlines()[-1]); print('C4FL (tile kernel)', '%.3f G/s'%(d['value']/1e9), '%.3f ms'%d['ms_per_step'])
except Exception as e: print('C4FL FAILED', e)
PY
