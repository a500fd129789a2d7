#!/bin/bash
# round 2, call p (1 GPU): ideal MHD 3-D on a 32 x 7 tile (8 warps, no register cap) against 32 x 8 (9 warps at 168 registers) / 32 x 6
mkdir -p gpurun_out
timeout 300 python tools/sweep_march.py M3 0,13,3 3 > gpurun_out/r02p_sweep_mhd_ty7.txt 2>&1
timeout 300 python tools/sweep_march.py M3r4 0,13,3 3 >> gpurun_out/r02p_sweep_mhd_ty7.txt 2>&1
cat gpurun_out/r02p_sweep_mhd_ty7.txt
