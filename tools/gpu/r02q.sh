#!/bin/bash
# round 2, call q (1 GPU): ideal MHD 3-D RK4 with / without the folded last stage on the 32 x 7 tile; the general marching configuration on 32 x 7 against the tile kernel
mkdir -p gpurun_out
timeout 300 python tools/sweep_march.py M3r4 0 3 > gpurun_out/r02q_sweep_mhd_fold.txt 2>&1
HB_RK_FOLD=1 timeout 300 python tools/sweep_march.py M3r4 0 3 >> gpurun_out/r02q_sweep_mhd_fold.txt 2>&1
cat gpurun_out/r02q_sweep_mhd_fold.txt
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
for name, extra in (("3D plm minmod + hll", dict(usePLM="plm cons", slopeLimiter="minmod", flux="hll")), ("3D plm van Leer + roe", dict(usePLM="plm cons", slopeLimiter="monotized central")),
                    ("3D no recon + roe (donor)", dict()), ("3D flux limiter superbee", dict(fluxLimiter="superbee"))):
    for sk, lab in ((2, "march(GEN)"), (1, "tile")):
        t(dict(base3, stage_kernel=sk, **extra), name + " [" + lab + "]")
PY
timeout 900 python /tmp/gen_time.py > gpurun_out/r02q_gen_vs_tile.txt 2>&1; cat gpurun_out/r02q_gen_vs_tile.txt
