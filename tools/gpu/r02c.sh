#!/bin/bash
# round 2, call c: per-stage marching configuration, halo-warp backoff; tests + sweeps (C4 RK4, C4r3, MHD 3-D)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1
tail -3 gpurun_out/r02c_pytest.log
timeout 300 python tools/sweep_march.py C4r3 0,1,10 3 > gpurun_out/r02c_sweep_c4r3.txt 2>&1
cat gpurun_out/r02c_sweep_c4r3.txt
timeout 300 python tools/sweep_march.py C4 0,1,5,10 3 > gpurun_out/r02c_sweep_c4.txt 2>&1
cat gpurun_out/r02c_sweep_c4.txt
HB_MARCH_PER_STAGE=0 timeout 300 python tools/sweep_march.py C4 1,5 3 > gpurun_out/r02c_sweep_c4_single.txt 2>&1
cat gpurun_out/r02c_sweep_c4_single.txt
timeout 300 python tools/sweep_march.py M3 0,3,10,12 3 > gpurun_out/r02c_sweep_m3.txt 2>&1
cat gpurun_out/r02c_sweep_m3.txt
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r02c_bench_c4.json 2> gpurun_out/r02c_bench_c4.err
cat gpurun_out/r02c_bench_c4.json | cut -c1-900
