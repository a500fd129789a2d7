#!/bin/bash
# round 2, call i (1 GPU): whole GPU suite, smoke, PAIR-variant sweep, every bench workload (double; float for C4 / C2), ncu launch list + full captures
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02i_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02i_smoke.txt 2>&1; tail -1 gpurun_out/r02i_smoke.txt
timeout 300 python tools/sweep_march.py C4 1,10,11 3 > gpurun_out/r02i_sweep_c4_pair.txt 2>&1; cat gpurun_out/r02i_sweep_c4_pair.txt
timeout 300 python tools/sweep_march.py C4r3 0,10 3 >> gpurun_out/r02i_sweep_c4_pair.txt 2>&1; tail -2 gpurun_out/r02i_sweep_c4_pair.txt
for w in C4 C4M C2 C3 C5; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r02i_bench_$w.json 2> gpurun_out/r02i_bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02i_bench_$w.json').read().strip().splitlines()[-1]); print('$w', '%.3f G/s'%(d['value']/1e9), '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'parity', (d.get('parity') or {}).get('rel_linf'), 'e2e %.3f'%(d['e2e']['value']/1e9), 'cpu %.2f M'%(d['cpu_baseline']['value']/1e6))
except Exception as e: print('$w', 'FAILED', e); print(open('gpurun_out/r02i_bench_$w.err').read()[-600:])
PY
done
for w in C4 C2; do
  timeout 600 python bench.py --workload $w --precision float --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench_${w}_f32.json 2> gpurun_out/r02i_bench_${w}_f32.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02i_bench_${w}_f32.json').read().strip().splitlines()[-1]); print('$w f32', '%.3f G/s'%(d['value']/1e9), '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'parity', (d.get('parity') or {}).get('rel_linf'))
except Exception as e: print('$w f32', 'FAILED', e); print(open('gpurun_out/r02i_bench_${w}_f32.err').read()[-600:])
PY
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02i_bench_reference_arm.json 2>&1; cut -c1-300 gpurun_out/r02i_bench_reference_arm.json
# ncu: launch list of the default bench command, then full captures of the stage kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02i_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r02i_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fv_march3 -s 8 -c 4 -f -o gpurun_out/r02i_march3_c4 python tools/sweep_march.py C4 1 1 > gpurun_out/r02i_ncu_c4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fv_march3 -s 3 -c 1 -f -o gpurun_out/r02i_march3_m3 python tools/sweep_march.py M3 0 1 > gpurun_out/r02i_ncu_m3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fv_march2d -s 3 -c 1 -f -o gpurun_out/r02i_march2d_c3 python tools/sweep_march.py C3 0 1 > gpurun_out/r02i_ncu_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fv_march2d -s 4 -c 1 -f -o gpurun_out/r02i_march2d_c2 python tools/sweep_march.py C2 0 1 > gpurun_out/r02i_ncu_c2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
