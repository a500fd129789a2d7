"""Sweep the marching-kernel tile configurations ($HB_MARCH_CFG) on one workload: per-launch stage-kernel time (CUDA events on the
library stream, eager launches) and the difference of the final state against the first configuration.  Development aid."""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hydrob200
from importlib import import_module
hb = import_module("hydro-cl-lua_b200._lib")

WORK = {
    "C4": (dict(eqn="euler", dim=3, gridSize=[512, 512, 128], mins=[-2] * 3, maxs=[2, 2, -1], initCond="sphere", usePLM="plm cons",
                slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.1), 640 / 4),
    "C4r3": (dict(eqn="euler", dim=3, gridSize=[512, 512, 128], mins=[-2] * 3, maxs=[2, 2, -1], initCond="sphere", usePLM="plm cons",
                  slopeLimiter="minmod", integrator="Runge-Kutta 3, TVD", cfl=.1), 640 / 4),
    "C2": (dict(eqn="euler", dim=2, gridSize=[2048, 2048], initCond="Kelvin-Helmholtz", usePLM="plm cons",
                slopeLimiter="minmod", integrator="Runge-Kutta 4, TVD", cfl=.15), 840 / 4),
    "C3": (dict(eqn="mhd", dim=2, gridSize=[2048, 2048], initCond="Orszag-Tang", usePLM="plm cons",
                slopeLimiter="minmod", integrator="Runge-Kutta 3, TVD", cfl=.15), 512 / 3),
    "M3r4": (dict(eqn="mhd", dim=3, gridSize=[256, 256, 64], initCond="Orszag-Tang", usePLM="plm cons", mins=[-2] * 3, maxs=[2] * 3,
                  slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.1), 1024 / 4),
    "M3": (dict(eqn="mhd", dim=3, gridSize=[256, 256, 64], initCond="Orszag-Tang", usePLM="plm cons", mins=[-2] * 3, maxs=[2] * 3,
                slopeLimiter="minmod", integrator="Runge-Kutta 3, TVD", cfl=.1), 512 / 3),
}
work = sys.argv[1]
cfgs = [int(x) for x in sys.argv[2].split(",")]
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg, bytes_per_cell_stage = WORK[work]
ref = None
for c in cfgs:
    os.environ["HB_MARCH_CFG"] = str(c)
    try:
        S = hydrob200.FiniteVolumeSolver(dict(cfg, use_graph=False))
    except Exception as e:
        print("cfg %d: %s" % (c, e), flush=True)
        continue
    B = S.backend
    desc = B.describe().split("\n")[0]
    S.update(1)
    hb.check(B.L.hb_fv_profile(B.h, 1))
    S.update(nsteps)
    ms, n = C.c_double(), C.c_longlong()
    hb.check(B.L.hb_fv_profile_read(B.h, C.byref(ms), C.byref(n)))
    hb.check(B.L.hb_fv_profile(B.h, 0))
    per = ms.value / n.value
    cells = int(np.prod(cfg["gridSize"]))
    U = S.interior()
    if ref is None:
        ref = U
        diff = 0.
    else:
        diff = max(np.abs(U[..., q] - ref[..., q]).max() / np.abs(ref[..., q]).max() for q in range(U.shape[-1]) if np.abs(ref[..., q]).max() > 0)
    print("cfg %2d  %.4f ms/stage  %.3f G cell-stages/s  %.0f GB/s alg  rel diff vs first %.2e  | %s" % (
        c, per, cells / per / 1e6, cells * bytes_per_cell_stage / per / 1e6, diff, desc), flush=True)
    del S, B
