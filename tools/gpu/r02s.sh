#!/bin/bash
# round 2, call s (4 GPUs): the shipped state on 2 and 4 ranks -- decomposed == single (bitwise; folded RK4 plan, TMA operands, rim / interior overlap), thin-slab weak and strong scaling at N = 1, 2, 4
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_slab_decomposition.py tests/test_gpu_fine_grained.py -m gpu -q > gpurun_out/r02s_pytest_4gpu.log 2>&1
tail -6 gpurun_out/r02s_pytest_4gpu.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { # name nproc args...
  name=$1; np=$2; shift 2
  if [ "$np" = 1 ]; then timeout 400 python bench.py --gpus 1 "$@" > gpurun_out/r02s_$name.json 2> gpurun_out/r02s_$name.err
  else timeout 400 $TR --master-port $((29500 + RANDOM % 400)) --nproc-per-node $np bench.py --gpus $np "$@" > gpurun_out/r02s_$name.json 2> gpurun_out/r02s_$name.err; fi
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02s_$name.json').read().strip().splitlines()[-1]); print('$name', 'N', d['n_gpus'], '%.3f G/s'%(d['value']/1e9), '%.3f ms'%d['ms_per_step'], d['config']['per_gpu_grid'], d['roofline']['kernel_config'].split('Lbufs')[-1], 'e2e/gpu %.3f'%(d['e2e'].get('per_gpu',0)/1e9), 'parity', (d.get('parity') or {}).get('rel_linf'))
except Exception as e: print('$name', 'FAILED', e); print(open('gpurun_out/r02s_$name.err').read()[-600:])
PY
}
run thin_n1 1 --steps 10 --warmup 3 --per-gpu-planes 64 --no-cpu-baseline --no-parity
run thin_n2 2 --steps 10 --warmup 3 --per-gpu-planes 64 --no-cpu-baseline
run thin_n4 4 --steps 10 --warmup 3 --per-gpu-planes 64 --no-cpu-baseline
run strong_n2 2 --steps 5 --warmup 3 --scaling strong --no-cpu-baseline
run strong_n4 4 --steps 5 --warmup 3 --scaling strong --no-cpu-baseline
run default_n2 2 --steps 10 --warmup 3
