#!/bin/bash
# round 2, call d: new parity tests (full-size sub-slab, every limiter, Alfven wave), bench with the in-run parity leg, C4M, host info
mkdir -p gpurun_out
(nproc; free -g; lscpu | head -20; nvidia-smi topo -m) > gpurun_out/r02d_host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02d_pytest.log 2>&1
tail -25 gpurun_out/r02d_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02d_bench_c4.json 2> gpurun_out/r02d_bench_c4.err
cut -c1-1500 gpurun_out/r02d_bench_c4.json; tail -3 gpurun_out/r02d_bench_c4.err
timeout 600 python bench.py --workload C4M --steps 5 --warmup 3 > gpurun_out/r02d_bench_c4m.json 2> gpurun_out/r02d_bench_c4m.err
cut -c1-1500 gpurun_out/r02d_bench_c4m.json; tail -3 gpurun_out/r02d_bench_c4m.err
