#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
run() { # workload cfg
  HB_MARCH_CFG=$2 timeout 300 python bench.py --workload $1 --no-cpu-baseline --steps 5 > gpurun_out/bench_$1_cfg$2.json 2> gpurun_out/bench_$1_cfg$2.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$1_cfg$2.json"))
    print("$1 cfg$2 value %.4g  stage_ms %.3f frac %.4f e2e %.4g t=%.6g" % (d["value"], d["roofline"]["stage_kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["config"]["t"]))
except Exception as e:
    print("$1 cfg$2 failed", e)
PY
}
for cfg in 0 1 5; do run C4 $cfg; done
for cfg in 0 1; do run C2 $cfg; done
run C3 0
HB_MARCH_CFG=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fv_march -s 5 -c 1 -o gpurun_out/fv_march_c4_cfg1 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --grid 512,512,128 > gpurun_out/ncu_full.log 2>&1

