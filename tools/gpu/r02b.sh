#!/bin/bash
# round 2, call b: leaner fv_march3 (term-list epilogue, lean operand staging, new Euler Roe core): tests + sweep + ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1
tail -5 gpurun_out/r02b_pytest.log
timeout 300 python tools/sweep_march.py C4r3 0,1,2,10 3 > gpurun_out/r02b_sweep_c4r3.txt 2>&1
cat gpurun_out/r02b_sweep_c4r3.txt
timeout 300 python tools/sweep_march.py C4 1,2,5,10 3 > gpurun_out/r02b_sweep_c4.txt 2>&1
cat gpurun_out/r02b_sweep_c4.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fv_march3 -s 5 -c 1 -f -o gpurun_out/r02b_march3_c4 python tools/sweep_march.py C4 1 1 > gpurun_out/r02b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fv_march3 -s 5 -c 1 -f -o gpurun_out/r02b_march3_c4r3 python tools/sweep_march.py C4r3 0 1 > gpurun_out/r02b_ncu2.log 2>&1
tail -2 gpurun_out/r02b_ncu.log gpurun_out/r02b_ncu2.log
