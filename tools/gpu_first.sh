#!/bin/bash
# First GPU pass: parity tests, bench, ncu launch list, one full capture of the stage kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench rc=$?" >> gpurun_out/bench_c4.err
timeout 300 python bench.py --workload C2 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 300 python bench.py --workload C3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fv_stage -s 4 -c 2 -o gpurun_out/fv_stage_c4 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_c4.json; tail -2 gpurun_out/bench_c4.err; cat gpurun_out/bench_c2.json gpurun_out/bench_c3.json
