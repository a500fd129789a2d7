#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; cat gpurun_out/bench_c4.json; tail -2 gpurun_out/bench_c4.err
timeout 300 python bench.py --workload C2 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
timeout 300 python bench.py --workload C3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fv_march -s 5 -c 1 -o gpurun_out/fv_march_c4 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fv_march -s 5 -c 1 -o gpurun_out/fv_march_c2 -f python bench.py --workload C2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fv_march -s 4 -c 1 -o gpurun_out/fv_march_c3 -f python bench.py --workload C3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c3.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
