#!/bin/bash
# round 2, first GPU call: full GPU test suite on the new 3-D marching kernel + configuration sweep + one ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1
tail -5 gpurun_out/r02a_pytest.log
timeout 300 python tools/sweep_march.py C4r3 0,1,2,4,10 3 > gpurun_out/r02a_sweep_c4r3.txt 2>&1
cat gpurun_out/r02a_sweep_c4r3.txt
timeout 300 python tools/sweep_march.py C4 1,2,5,10 3 > gpurun_out/r02a_sweep_c4.txt 2>&1
cat gpurun_out/r02a_sweep_c4.txt
timeout 300 python tools/sweep_march.py M3 3,10,12 3 > gpurun_out/r02a_sweep_m3.txt 2>&1
cat gpurun_out/r02a_sweep_m3.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fv_march3 -s 5 -c 2 -f -o gpurun_out/r02a_march3_c4 python tools/sweep_march.py C4 1 1 > gpurun_out/r02a_ncu.log 2>&1
tail -3 gpurun_out/r02a_ncu.log
