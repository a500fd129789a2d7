#!/bin/bash
# round 2, call e (2 GPUs): rim / interior overlap, upgraded 2-D kernel: tests incl. the two-rank cases; thin-slab weak scaling 1 -> 2; strong scaling; 2-D sweeps
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m "gpu and not slow" -x -q > gpurun_out/r02e_pytest_2gpu.log 2>&1
tail -4 gpurun_out/r02e_pytest_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --per-gpu-planes 64 --no-cpu-baseline --no-parity > gpurun_out/r02e_thin_n1.json 2> gpurun_out/r02e_thin_n1.err
timeout 300 $TR --nproc-per-node 2 bench.py --gpus 2 --steps 10 --warmup 3 --per-gpu-planes 64 --no-cpu-baseline > gpurun_out/r02e_thin_n2.json 2> gpurun_out/r02e_thin_n2.err
HB_OVERLAP=0 timeout 300 $TR --nproc-per-node 2 bench.py --gpus 2 --steps 10 --warmup 3 --per-gpu-planes 64 --no-cpu-baseline > gpurun_out/r02e_thin_n2_instream.json 2> gpurun_out/r02e_thin_n2_instream.err
timeout 300 $TR --nproc-per-node 2 bench.py --gpus 2 --steps 5 --warmup 3 --scaling strong --no-cpu-baseline > gpurun_out/r02e_strong_n2.json 2> gpurun_out/r02e_strong_n2.err
for f in thin_n1 thin_n2 thin_n2_instream strong_n2; do python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02e_$f.json')); print('$f', d['value']/1e9, 'G/s', d['ms_per_step'], 'ms', d['config']['per_gpu_grid'], d['roofline']['kernel_config'][-60:], 'e2e', d['e2e']['value']/1e9)
except Exception as e: print('$f', 'FAILED', e); print(open('gpurun_out/r02e_$f.err').read()[-800:])
PY
done
timeout 300 python tools/sweep_march.py C2 0,1 3 > gpurun_out/r02e_sweep_c2.txt 2>&1; cat gpurun_out/r02e_sweep_c2.txt
timeout 300 python tools/sweep_march.py C3 0,1 3 > gpurun_out/r02e_sweep_c3.txt 2>&1; cat gpurun_out/r02e_sweep_c3.txt
