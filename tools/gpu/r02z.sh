#!/bin/bash
# round 2, call z (8 GPUs): thin-slab weak and strong scaling at N = 8 with the shipped kernels
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { name=$1; np=$2; shift 2
  timeout 70 $TR --master-port $((29500 + RANDOM % 400)) --nproc-per-node $np bench.py --gpus $np "$@" > gpurun_out/r02z_$name.json 2> gpurun_out/r02z_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02z_$name.json').read().strip().splitlines()[-1]); print('$name', 'N', d['n_gpus'], '%.3f G/s'%(d['value']/1e9), '%.3f ms'%d['ms_per_step'], d['config']['per_gpu_grid'])
except Exception as e: print('$name', 'FAILED', e); print(open('gpurun_out/r02z_$name.err').read()[-400:])
PY
}
run thin_n8 8 --steps 10 --warmup 3 --per-gpu-planes 64 --no-cpu-baseline --e2e-steps 1
run strong_n8 8 --steps 5 --warmup 3 --scaling strong --no-cpu-baseline --e2e-steps 1
