#!/bin/bash
# marching kernel: parity, then bench of the tile configurations
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for cfg in 0 1 2; do
  HB_MARCH_CFG=$cfg timeout 300 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/bench_c4_cfg$cfg.json 2> gpurun_out/bench_c4_cfg$cfg.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_c4_cfg$cfg.json"))
    print("C4 cfg$cfg value %.4g  stage_ms %.3f frac %.4f e2e %.4g" % (d["value"], d["roofline"]["stage_kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"]))
except Exception as e:
    print("C4 cfg$cfg failed", e)
PY
  tail -2 gpurun_out/bench_c4_cfg$cfg.err
done
for w in C2 C3; do for cfg in 0 1; do
  HB_MARCH_CFG=$cfg timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 5 > gpurun_out/bench_${w}_cfg$cfg.json 2> gpurun_out/bench_${w}_cfg$cfg.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${w}_cfg$cfg.json"))
    print("$w cfg$cfg value %.4g  stage_ms %.3f frac %.4f" % (d["value"], d["roofline"]["stage_kernel_ms"], d["roofline"]["frac"]))
except Exception as e:
    print("$w cfg$cfg failed", e)
PY
done; done
HB_MARCH_CFG=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fv_march -s 4 -c 1 -o gpurun_out/fv_march_c4 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --grid 512,512,128 > gpurun_out/ncu_full.log 2>&1
