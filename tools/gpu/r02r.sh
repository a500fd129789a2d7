#!/bin/bash
# round 2, call r (1 GPU): the shipped state -- whole GPU suite, smoke, every bench workload, reference arm, ncu launch list + full captures, PCIe ceiling
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02r_pytest_gpu.log 2>&1
tail -6 gpurun_out/r02r_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02r_smoke.txt 2>&1; tail -1 gpurun_out/r02r_smoke.txt
python tools/pcie_bw.py > gpurun_out/r02r_pcie.txt 2>&1; cat gpurun_out/r02r_pcie.txt
for w in C4 C4M C2 C3 C5; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r02r_bench_$w.json 2> gpurun_out/r02r_bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02r_bench_$w.json').read().strip().splitlines()[-1]); print('$w', '%.3f G/s'%(d['value']/1e9), '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'stage %.3f ms'%d['roofline']['stage_kernel_ms'], 'parity', (d.get('parity') or {}).get('rel_linf'), 'e2e %.3f'%(d['e2e']['value']/1e9), 'cpu %.2f M'%(d['cpu_baseline']['value']/1e6))
except Exception as e: print('$w', 'FAILED', e); print(open('gpurun_out/r02r_bench_$w.err').read()[-600:])
PY
done
for w in C4 C2; do
  timeout 600 python bench.py --workload $w --precision float --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02r_bench_${w}_f32.json 2> gpurun_out/r02r_bench_${w}_f32.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02r_bench_${w}_f32.json').read().strip().splitlines()[-1]); print('$w f32', '%.3f G/s'%(d['value']/1e9), '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'parity', (d.get('parity') or {}).get('rel_linf'))
except Exception as e: print('$w f32', 'FAILED', e); print(open('gpurun_out/r02r_bench_${w}_f32.err').read()[-600:])
PY
done
timeout 600 python bench.py > gpurun_out/r02r_bench_default.json 2> gpurun_out/r02r_bench_default.err; cut -c1-200 gpurun_out/r02r_bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02r_bench_reference_arm.json 2>&1; cut -c1-200 gpurun_out/r02r_bench_reference_arm.json
# ncu: launch list of the default bench command, then full captures of the stage kernels (exports only: the reports exceed the copy-back limit)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02r_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r02r_ncu_launches.log 2>&1
cap() { # name regex skip count cmd...
  name=$1; re=$2; sk=$3; cnt=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$re -s $sk -c $cnt -f -o /tmp/$name "$@" > gpurun_out/r02r_ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/r02r_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/r02r_${name}_source.csv 2>/dev/null
  ls -la /tmp/$name.ncu-rep
}
cap march3_c4 fv_march3 4 4 python tools/sweep_march.py C4 0 1
cap march3_m3 fv_march3 3 1 python tools/sweep_march.py M3 0 1
du -sh gpurun_out
