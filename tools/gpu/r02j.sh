#!/bin/bash
# round 2, call j (1 GPU): whole GPU suite incl. the general marching configurations, a flux-limiter bench line, every bench workload,
# ncu launch list + full captures exported to text ON THE BOX (the .ncu-rep files exceed gpurun_out's 64 MiB)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02j_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02j_smoke.txt 2>&1; tail -1 gpurun_out/r02j_smoke.txt
for w in C4 C4M C2 C3 C5; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r02j_bench_$w.json 2> gpurun_out/r02j_bench_$w.err
done
for w in C4 C2; do
  timeout 600 python bench.py --workload $w --precision float --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench_${w}_f32.json 2> gpurun_out/r02j_bench_${w}_f32.err
done
timeout 600 python bench.py --workload C4FL --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_bench_C4FL.json 2> gpurun_out/r02j_bench_C4FL.err
HB_MARCH_GEN=0 timeout 600 python bench.py --workload C4FL --steps 3 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02j_bench_C4FL_tile.json 2> gpurun_out/r02j_bench_C4FL_tile.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02j_bench_reference_arm.json 2> gpurun_out/r02j_bench_reference_arm.err
for f in C4 C4M C2 C3 C5 C4_f32 C2_f32 C4FL C4FL_tile; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02j_bench_$f.json').read().strip().splitlines()[-1]); print('$f', '%.3f G/s'%(d['value']/1e9), '%.3f ms'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], 'parity', (d.get('parity') or {}).get('rel_linf'), 'e2e %.3f'%(d['e2e']['value']/1e9), d['roofline']['kernel_config'][:60])
except Exception as e: print('$f', 'FAILED', e); print(open('gpurun_out/r02j_bench_$f.err').read()[-500:])
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02j_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r02j_ncu_launches.log 2>&1
cap() { # name regex skip count cmd...
  name=$1; re=$2; sk=$3; cnt=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$re -s $sk -c $cnt -f -o /tmp/$name "$@" > gpurun_out/r02j_ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/r02j_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/r02j_${name}_source.csv 2>/dev/null
  ls -la /tmp/$name.ncu-rep
}
cap march3_c4 fv_march3 8 4 python tools/sweep_march.py C4 0 1
cap march3_m3 fv_march3 3 1 python tools/sweep_march.py M3 0 1
cap march2d_c3 fv_march2d 3 1 python tools/sweep_march.py C3 0 1
cap march2d_c2 fv_march2d 4 1 python tools/sweep_march.py C2 0 1
du -sh gpurun_out
