#!/bin/bash
# round 2, call k (1 GPU): whole suite after the GEN fixes, ncu of the C4 stage kernels and of the ADM stage (traffic), GEN vs tile timing for PLM + HLL
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02k_pytest_gpu.log
cat > /tmp/gen_time.py <<'PY'
import os, sys, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, hydrob200
from importlib import import_module
hb = import_module("hydro-cl-lua_b200._lib")
def t(cfg, label):
    S = hydrob200.FiniteVolumeSolver(dict(cfg, use_graph=False)); B = S.backend
    S.update(1); hb.check(B.L.hb_fv_profile(B.h, 1)); S.update(3)
    ms, n = C.c_double(), C.c_longlong(); hb.check(B.L.hb_fv_profile_read(B.h, C.byref(ms), C.byref(n))); hb.check(B.L.hb_fv_profile(B.h, 0))
    cells = int(np.prod(cfg["gridSize"]))
    print("%-34s %8.3f ms/stage  %6.2f G cell-stages/s | %s" % (label, ms.value / n.value, cells / (ms.value / n.value) / 1e6, B.describe().split("\n")[0][:110]), flush=True)
base3 = dict(eqn="euler", dim=3, gridSize=[256, 256, 256], mins=[-2]*3, maxs=[2]*3, initCond="sphere", integrator="Runge-Kutta 4", cfl=.1)
base2 = dict(eqn="euler", dim=2, gridSize=[2048, 2048], initCond="Kelvin-Helmholtz", integrator="Runge-Kutta 4, TVD", cfl=.15)
for name, extra in (("3D plm minmod + hll", dict(usePLM="plm cons", slopeLimiter="minmod", flux="hll")), ("3D plm van Leer + roe", dict(usePLM="plm cons", slopeLimiter="monotized central")),
                    ("3D no recon + roe (donor)", dict()), ("3D flux limiter superbee", dict(fluxLimiter="superbee"))):
    for sk, lab in ((0, "default"), (1, "tile")):
        if name.startswith("3D flux") and sk == 0: sk, lab = 2, "march(GEN)"
        t(dict(base3, stage_kernel=sk, **extra), name + " [" + lab + "]")
for name, extra in (("2D plm minmod + hll", dict(usePLM="plm cons", slopeLimiter="minmod", flux="hll")), ("2D flux limiter superbee", dict(fluxLimiter="superbee"))):
    for sk, lab in ((0, "default"), (1, "tile")):
        if name.startswith("2D flux") and sk == 0: sk, lab = 2, "march(GEN)"
        t(dict(base2, stage_kernel=sk, **extra), name + " [" + lab + "]")
PY
timeout 900 python /tmp/gen_time.py > gpurun_out/r02k_gen_vs_tile.txt 2>&1; cat gpurun_out/r02k_gen_vs_tile.txt
cap() { # name regex skip count cmd...
  name=$1; re=$2; sk=$3; cnt=$4; shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$re -s $sk -c $cnt -f -o /tmp/$name "$@" > gpurun_out/r02k_ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/r02k_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/r02k_${name}_source.csv 2>/dev/null
  ls -la /tmp/$name.ncu-rep
}
cap march3_c4 fv_march3 4 4 python tools/sweep_march.py C4 0 1
timeout 900 ncu --set full --clock-control none -k regex:adm_ -s 10 -c 10 -f -o /tmp/adm_c5 python bench.py --workload C5 --grid 128,128,128 --steps 1 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r02k_ncu_adm.log 2>&1
ncu -i /tmp/adm_c5.ncu-rep --page raw --csv > gpurun_out/r02k_adm_c5_128_raw.csv 2>/dev/null
du -sh gpurun_out
