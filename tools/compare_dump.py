#!/usr/bin/env python
"""Compare a state dump written by the reference (GridSolver:save / saveOnExit -> <prefix>_UBuf.fits, hydro/solver/gridsolver.lua:1410-1470)
with the state this library reaches for the same configuration -- the hook SURVEY 8c asks for, so that
parity can be pinned on the reference's own OpenCL output wherever the reference can be run.

    python tools/compare_dump.py --fits run_UBuf.fits --cfg '{"eqn": "euler", "dim": 2, "gridSize": [64, 64], ...}' --steps 20
    python tools/compare_dump.py --fits run_UBuf.fits --case C2_kh_rk4tvd_minmod --steps 20 [--until-t 0.1]

The configuration keys are the reference's config.lua names (eqn, dim, gridSize, mins, maxs, boundary, initCond, usePLM, slopeLimiter,
fluxLimiter, flux, integrator, cfl, fixedDT, useCTU, useGravity, noDiv).  Prints the relative L-infinity difference per state variable
(ghost cells excluded unless --ghosts) and exits 1 if any exceeds --tol (default 1e-12, the north-star bar in double)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--fits", required=True)
    ap.add_argument("--cfg", help="JSON object of solver arguments")
    ap.add_argument("--case", help="a named case of tests/cases.py instead of --cfg")
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--until-t", type=float, default=None, help="run update() until t >= this (hydro/app.lua:1590 exit rule)")
    ap.add_argument("--ghosts", action="store_true", help="compare ghost cells too")
    ap.add_argument("--tol", type=float, default=1e-12)
    args = ap.parse_args(argv)
    import hydrob200
    if args.case:
        from cases import ADM_CASES, CASES
        cfg, n = dict(CASES, **ADM_CASES)[args.case]
        cfg = dict(cfg)
        if args.steps is None and args.until_t is None:
            args.steps = n
    else:
        cfg = json.loads(args.cfg)
    S = hydrob200.FiniteVolumeSolver(cfg)          # the CUDA library: no other backend here
    if args.until_t is not None:
        while S.t < args.until_t:
            S.update()
    else:
        for _ in range(args.steps or 0):
            S.update()
    return compare(S, args.fits, args.ghosts, args.tol)


def compare(S, fits, ghosts=False, tol=1e-12):
    """S: a solver advanced to the dump's time; -> 0 within tol, 1 beyond, 2 shape mismatch."""
    class A:
        pass
    args = A()
    args.fits, args.ghosts, args.tol = fits, ghosts, tol
    ref = S.loadBuffer(args.fits)
    got = S.getState()
    if ref.shape != got.shape:
        print("dump shape %s does not match the solver's state %s" % (ref.shape, got.shape))
        return 2
    if not args.ghosts:
        g = S.numGhost
        sl = tuple(slice(g, -g) if S.dim > 2 - a else slice(None) for a in range(3))
        ref, got = ref[sl], got[sl]
    worst = 0.
    names = list(getattr(S.eqn, "consVars", [])) or [str(q) for q in range(got.shape[-1])]
    for q in range(got.shape[-1]):
        scale = np.abs(ref[..., q]).max()
        err = np.abs(got[..., q] - ref[..., q]).max()
        rel = err / scale if scale > 0 else err
        worst = max(worst, rel)
        print("%-12s rel Linf %.3e   (max |ref| %.6g)" % (names[q] if q < len(names) else q, rel, scale))
    print("t = %.15g   worst %.3e   %s" % (S.t, worst, "OK" if worst <= args.tol else "EXCEEDS --tol %g" % args.tol))
    return 0 if worst <= args.tol else 1


if __name__ == "__main__":
    sys.exit(main())
