#!/bin/bash
# round 2, call n (1 GPU): two CTAs per SM (March3Cfg::MINB, half-height tiles) against one 16-warp CTA
mkdir -p gpurun_out
timeout 300 python tools/sweep_march.py C4r3 0,13,14,10 3 > gpurun_out/r02n_sweep_minb2.txt 2>&1; cat gpurun_out/r02n_sweep_minb2.txt
timeout 300 python tools/sweep_march.py C4 0,13,14 3 >> gpurun_out/r02n_sweep_minb2.txt 2>&1; tail -3 gpurun_out/r02n_sweep_minb2.txt
