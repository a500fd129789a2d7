#!/bin/bash
# ncu evidence for profiles/r01c_*: launch list of the default bench command + one --set full capture per dominant kernel
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fv_march -s 5 -c 1 -o gpurun_out/fv_march_c4 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fv_march -s 5 -c 1 -o gpurun_out/fv_march_c2 -f python bench.py --workload C2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fv_march -s 4 -c 1 -o gpurun_out/fv_march_c3 -f python bench.py --workload C3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c3.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:adm_ -s 12 -c 6 -o gpurun_out/adm_c5 -f python bench.py --workload C5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5_256.csv python bench.py --workload C5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_c5.log 2>&1
ls -la gpurun_out/*.ncu-rep
# summaries on the box (the reports together exceed what gpurun copies back)
declare -A CELLS=([c4]=134217728 [c2]=4194304 [c3]=16777216)
for w in c4 c2 c3; do
  python tools/ncu_summary.py full gpurun_out/fv_march_$w.ncu-rep ${CELLS[$w]} > gpurun_out/fv_march_${w}_full.txt 2>&1
  python tools/ncu_source_summary.py gpurun_out/fv_march_$w.ncu-rep > gpurun_out/fv_march_${w}_source.txt 2>&1
done
ncu -i gpurun_out/adm_c5.ncu-rep --page raw --csv > gpurun_out/adm_c5_raw.csv 2>/dev/null
python tools/ncu_summary.py launches gpurun_out/launches_c4.csv > gpurun_out/launches_c4_summary.txt 2>&1
python tools/ncu_summary.py launches gpurun_out/launches_c5_256.csv > gpurun_out/launches_c5_256_summary.txt 2>&1
rm -f gpurun_out/fv_march_c2.ncu-rep gpurun_out/fv_march_c3.ncu-rep gpurun_out/adm_c5.ncu-rep
ls -la gpurun_out
