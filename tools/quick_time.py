"""Quick device timing of whole updates for a few shapes (development aid; bench.py is the contract)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import hydrob200

def timeit(name, cfg, nsteps, balg):
    S = hydrob200.FiniteVolumeSolver(cfg)
    print(S.backend.describe())
    S.update(2)
    ctx = S.backend.ctx
    ctx.sync()
    ctx.timerStart()
    S.update(nsteps)
    ms = ctx.timerStop()
    cells = 1
    for n in S.sizeWithoutBorder: cells *= n
    cups = cells * nsteps / (ms * 1e-3)
    print("%s: %.3f ms/update, %.3f G cell-updates/s, %.1f GB/s algorithmic (%.1f%% of 6551.7), t=%g dt=%g" % (
        name, ms / nsteps, cups / 1e9, cups * balg / 1e9, 100 * cups * balg / 6551.7e9, S.t, S.dt), flush=True)

which = sys.argv[1:] or ["C4_256", "C2", "C3_2048"]
cfgs = {
    "C4_128": (dict(eqn="euler", dim=3, gridSize=[128] * 3, mins=[-2] * 3, maxs=[2] * 3, initCond="sphere", usePLM="plm cons",
                    slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.1), 5, 640),
    "C4_256": (dict(eqn="euler", dim=3, gridSize=[256] * 3, mins=[-2] * 3, maxs=[2] * 3, initCond="sphere", usePLM="plm cons",
                    slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.1), 5, 640),
    "C4_512": (dict(eqn="euler", dim=3, gridSize=[512] * 3, mins=[-2] * 3, maxs=[2] * 3, initCond="sphere", usePLM="plm cons",
                    slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.1), 3, 640),
    "C2": (dict(eqn="euler", dim=2, gridSize=[2048, 2048], initCond="Kelvin-Helmholtz", usePLM="plm cons",
                slopeLimiter="minmod", integrator="Runge-Kutta 4, TVD", cfl=.15), 10, 840),
    "C3_2048": (dict(eqn="mhd", dim=2, gridSize=[2048, 2048], initCond="Orszag-Tang", usePLM="plm cons",
                     slopeLimiter="minmod", integrator="Runge-Kutta 3, TVD", cfl=.15), 10, 512),
    "C3": (dict(eqn="mhd", dim=2, gridSize=[4096, 4096], initCond="Orszag-Tang", usePLM="plm cons",
                slopeLimiter="minmod", integrator="Runge-Kutta 3, TVD", cfl=.15), 5, 512),
    # the ops -- NoDiv (20 Jacobi sweeps after every step) and self-gravity (20 sweeps inside every stage; tile kernel)
    "C3_nodiv": (dict(eqn="mhd", dim=2, gridSize=[4096, 4096], initCond="Orszag-Tang", usePLM="plm cons",
                      slopeLimiter="minmod", integrator="Runge-Kutta 3, TVD", cfl=.15, noDiv="jacobi"), 5, 512),
    "C3_2048_nodiv": (dict(eqn="mhd", dim=2, gridSize=[2048, 2048], initCond="Orszag-Tang", usePLM="plm cons",
                           slopeLimiter="minmod", integrator="Runge-Kutta 3, TVD", cfl=.15, noDiv="jacobi"), 5, 512),
    "C4_256_grav": (dict(eqn="euler", dim=3, gridSize=[256] * 3, mins=[-2] * 3, maxs=[2] * 3, initCond="sphere", usePLM="plm cons",
                         slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.1, useGravity=True), 3, 640),
    "C4_256_tile": (dict(eqn="euler", dim=3, gridSize=[256] * 3, mins=[-2] * 3, maxs=[2] * 3, initCond="sphere", usePLM="plm cons",
                         slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.1, stage_kernel=1), 3, 640),
    "C2_grav": (dict(eqn="euler", dim=2, gridSize=[2048, 2048], initCond="Kelvin-Helmholtz", usePLM="plm cons",
                     slopeLimiter="minmod", integrator="Runge-Kutta 4, TVD", cfl=.15, useGravity=True), 5, 840),
    # useCTU -- the reference's unfused kernel sequence (hb_ctu_kernels.cuh)
    "C2_ctu": (dict(eqn="euler", dim=2, gridSize=[2048, 2048], initCond="Kelvin-Helmholtz", usePLM="plm cons",
                    slopeLimiter="minmod", integrator="Runge-Kutta 4, TVD", cfl=.15, useCTU=True), 5, 840),
    "C4_256_ctu": (dict(eqn="euler", dim=3, gridSize=[256] * 3, mins=[-2] * 3, maxs=[2] * 3, initCond="sphere", usePLM="plm cons",
                        slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.1, useCTU=True), 3, 640),
}
for w in which:
    cfg, n, balg = cfgs[w]
    timeit(w, cfg, n, balg)
