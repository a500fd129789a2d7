#!/bin/bash
# round 2, call w (1 GPU): PAIR (x and y flux cores as one block) with TMA operands on 32 x 7 / 11 / 15 tiles, against the default, Euler RK4 512 x 512 x 128
mkdir -p gpurun_out
timeout 400 python tools/sweep_march.py C4 0,14,15,16,2,1 3 > gpurun_out/r02w_sweep_pair.txt 2>&1; cat gpurun_out/r02w_sweep_pair.txt
