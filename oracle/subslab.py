"""TEST INFRASTRUCTURE (oracle/): parity of a FULL-SIZE CUDA run against the CPU oracle on a sub-slab.

The oracle keeps every buffer of the reference (UBuf, ULRBuf, fluxBuf, the integrator's U^k / L^k: ~0.9 KB per cell) and advances
7-10 M cell-updates/s on the GPU box's host cores, so a BASELINE-size 3-D grid (512^3: 137 M cells) is out of its reach, in memory
and in time.  What is compared instead: `planes` interior planes of the slowest axis, after `nsteps` updates of the WHOLE grid on the
GPU, against the oracle advancing only those planes plus the margin their domain of dependence needs (2 cells per stage and side:
the stencil radius of 'plm cons' + flux), started from the GPU solver's own initial state and driven with the GPU run's own dt
sequence (the CFL minimum is over the whole grid, which the sub-slab does not see).  The oracle's boundary condition at the cut
contaminates exactly the margin, which is discarded.  x-y extent, tile geometry, ring and chunk structure of the marching kernel are
the full-size ones.  If the margins reach the physical faces the comparison is against the full grid's oracle.

Used by tests/test_gpu_fullsize.py and by bench.py's `parity` leg (as the checker; nothing here is timed as the product)."""
import numpy as np


def rel_linf_grouped(a, b):
    """relative L-infinity per variable group (the components of a vector field share one scale)"""
    nv = a.shape[-1]
    groups = [[0], [1, 2, 3], [4]] + ([[5, 6, 7]] if nv >= 8 else [])
    out = []
    for g in groups:
        scale = np.abs(b[..., g]).max()
        err = np.abs(a[..., g] - b[..., g]).max()
        out.append(float(err / scale) if scale > 0 else float(err))
    return max(out), out


def subslab_parity(hydrob200, oracle, cfg, nsteps, planes, G=None, gpu_kwargs=None, nthreads=0):
    """-> dict(rel_linf, per_group, steps, planes, oracle_planes, bit_identical, dts).  `G`: an already constructed GPU solver in its
    initial state (bench.py hands its own); it is left `nsteps` steps further.  `G` may be a list of solvers in the same initial state
    (strict and production builds): all of them and the oracle are stepped with the FIRST one's dt sequence, and the result carries one
    entry per solver under "solvers"."""
    if G is None:
        G = hydrob200.FiniteVolumeSolver(dict(cfg, **(gpu_kwargs or {})))
    others = []
    if isinstance(G, (list, tuple)):
        G, others = G[0], list(G[1:])
    dim = int(cfg["dim"])
    ax = dim - 1
    g = G.numGhost
    N = G.sizeWithoutBorder[ax]
    nI = G.eqn.numIntStates
    stages = max(1, int(G.rkOrder))
    margin = 2 * stages * nsteps
    planes = min(planes, N)
    k0 = (N - planes) // 2
    k1 = k0 + planes
    lo, hi = max(0, k0 - margin), min(N, k1 + margin)
    if lo > 0 and hi < N:
        pass
    else:
        lo, hi = 0, N                      # the margins reach a physical face: take the whole axis (its boundary conditions are the real ones)
    dx = (G.maxs[ax] - G.mins[ax]) / float(N)
    sub = dict(cfg)
    gs = list(cfg["gridSize"]); gs[ax] = hi - lo
    mins = list(G.mins); maxs = list(G.maxs)
    mins[ax] = G.mins[ax] + lo * dx
    maxs[ax] = G.mins[ax] + hi * dx
    if (maxs[ax] - mins[ax]) / float(hi - lo) != dx:
        raise ValueError("sub-slab cell size differs from the grid's in the last bit: choose a dyadic domain / plane count")
    sub.update(gridSize=gs, mins=mins, maxs=maxs, backend=oracle.OracleBackendThreads(nthreads) if nthreads else oracle.OracleBackend)
    if lo > 0:
        # interior cut: whatever the boundary method does there is discarded with the margin; freeflow is the cheapest
        b = dict(sub.get("boundary") or {})
        names = ("x", "y", "z")
        for f in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")[:2 * dim]:
            b.setdefault(f, G.boundaryMethods[f])
        b[names[ax] + "min"] = "freeflow"; b[names[ax] + "max"] = "freeflow"
        sub["boundary"] = b
    U0 = G.getState()                                           # [Sz, Sy, Sx, nS], ghost cells included
    sl = [slice(None)] * 3
    sl[2 - ax] = slice(lo, hi + 2 * g)
    O = hydrob200.FiniteVolumeSolver(sub)
    O.setState(np.ascontiguousarray(U0[sl[0], sl[1], sl[2]]))
    del U0
    dts = []
    for _ in range(nsteps):
        dt = G.calcDT()
        dts.append(dt)
        G.step(dt)
        for H in others:
            H.step(dt)
        O.step(dt)
    sg = [slice(g, -g) if j < dim else slice(None) for j in (2, 1, 0)]
    so = list(sg)
    sg[2 - ax] = slice(g + k0, g + k1)
    so[2 - ax] = slice(g + k0 - lo, g + k1 - lo)
    b = O.getState()[so[0], so[1], so[2]][..., :nI]
    solvers = []
    for H in [G] + others:
        a = H.getState()[sg[0], sg[1], sg[2]][..., :nI]
        err, per = rel_linf_grouped(a, b)
        solvers.append(dict(rel_linf=err, per_group=per, finite=bool(np.isfinite(a).all()), bit_identical=bool(np.array_equal(a, b))))
    a = G.getState()[sg[0], sg[1], sg[2]][..., :nI]
    err, per = rel_linf_grouped(a, b)
    return dict(solvers=solvers, rel_linf=err, per_group=per, steps=nsteps, planes=planes, oracle_planes=hi - lo, finite=bool(np.isfinite(a).all()),
                bit_identical=bool(np.array_equal(a, b)), dts=dts, against="oracle (CPU restatement) on %d planes of the slowest axis, "
                "GPU dt sequence" % (hi - lo))
