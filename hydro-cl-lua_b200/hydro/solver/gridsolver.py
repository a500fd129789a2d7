"""GridSolver: structured grid, ghost cells, boundary methods (hydro/solver/gridsolver.lua).

Reference: numGhost = 2 :41; gridSize += 2*numGhost, unused dims size 1 :94-95; mins/maxs default +-1 :75-76;
grid_dx for all three axes :406-409; cell positions hydro/coord/coord.lua:1399-1458;
boundaryMethods default 'freeflow' :121-135, classes :638-780; usePLM/slopeLimiter :105-120 (PLM requires
fluxLimiter == 'donor cell', :119); calcExactError :1337-1366.
"""
import numpy as np

from .. import app as hydro_app
from .solverbase import SolverBase

boundaryIds = {"periodic": 0, "mirror": 1, "freeflow": 2, "none": 3, "linear": 4, "quadratic": 5, "fixed": 6}   # gridsolver.lua:638-846
xNames = ("x", "y", "z")
minmaxs = ("min", "max")
# 'plm athena': plm.cl:782-879 as the reference tree has it (result->L = cons(Wrv), result->R = cons(Wlv), :877-878);
# 'plm athena, recorded face order': L = left, R = right face -- reproduces the errors recorded in tests/test-order/schemes.lua
# 'plm prim': plm.cl:191-253; 'piecewise constant': plm.cl:10-24 (L = R = U: the same fluxes as no PLM with the donor-cell flux limiter)
# 'plm eig': plm.cl:256-427; 'plm eig prim' / 'plm eig prim ref': plm.cl:536-778 (", other face order": L and R exchanged)
plmIds = {None: 0, False: 0, "plm cons": 1, "plm athena": 2, "plm athena, recorded face order": 3, "plm prim": 4, "plm cons with flux": 5,
          "piecewise constant": 0, "plm eig": 6, "plm eig prim": 7, "plm eig prim ref": 8,
          "plm eig prim, other face order": 9, "plm eig prim ref, other face order": 10}


class GridSolver(SolverBase):
    name = "GridSolver"
    numGhost = 2

    def initMeshVars(self, args):
        dim = self.dim
        gs = args.get("gridSize", [256, 1, 1])
        if isinstance(gs, (int, float)):
            gs = [gs]
        gs = list(gs) + [1] * (3 - len(gs))
        self.sizeWithoutBorder = [int(gs[i]) if i < dim else 1 for i in range(3)]
        g = self.numGhost
        self.gridSize = [n + 2 * g if i < dim else 1 for i, n in enumerate(self.sizeWithoutBorder)]
        self.numCells = self.gridSize[0] * self.gridSize[1] * self.gridSize[2]
        self.stepsize = [1, self.gridSize[0], self.gridSize[0] * self.gridSize[1]]
        self.mins = [float(v) for v in args.get("mins", (-1., -1., -1.))]
        self.maxs = [float(v) for v in args.get("maxs", (1., 1., 1.))]
        # slab decomposition (hydro/solver/choppedup.py): this process owns localSizeWithoutBorder cells at localOffset
        self.comm = args.get("comm")
        if self.comm is not None:
            self.localSizeWithoutBorder, self.localOffset = self.comm.split(self.sizeWithoutBorder, dim)
        else:
            self.localSizeWithoutBorder, self.localOffset = list(self.sizeWithoutBorder), [0, 0, 0]
        self.localGridSize = [n + 2 * g if i < dim else 1 for i, n in enumerate(self.localSizeWithoutBorder)]
        self.localNumCells = self.localGridSize[0] * self.localGridSize[1] * self.localGridSize[2]

    def initObjs(self, args):
        super().initObjs(args)
        # initCond-supplied domain (solverbase.lua:1546-1560).  DELIBERATE DEVIATION: the reference lets the initCond
        # overwrite cfg mins / maxs / boundary unconditionally; here an explicit cfg entry wins, so that a parity case can put any
        # initial condition on any domain and boundary set.  Without the cfg entry the behaviour is the reference's.
        if self.initCond.mins is not None and "mins" not in args:
            self.mins = [float(v) for v in self.initCond.mins]
        if self.initCond.maxs is not None and "maxs" not in args:
            self.maxs = [float(v) for v in self.initCond.maxs]
        self.initCondMins = [float(v) for v in args.get("initCondMins", self.mins)]
        self.initCondMaxs = [float(v) for v in args.get("initCondMaxs", self.maxs)]
        self.grid_dx = [(self.maxs[j] - self.mins[j]) / float(self.sizeWithoutBorder[j]) for j in range(3)]
        self.mindx = min(self.grid_dx)
        self.usePLM = args.get("usePLM") or None
        if self.usePLM not in plmIds:
            raise NotImplementedError("usePLM=%r: built are %s ('ppm' is not)" % (self.usePLM, sorted(k for k in plmIds if isinstance(k, str))))
        self.plmId = plmIds[self.usePLM]
        # gridsolver.lua:106: `limiterNames:find(args.slopeLimiter) or 1` -- an omitted slopeLimiter is 'donor cell' (zero slope)
        self.slopeLimiter = hydro_app.limiterIndex(args.get("slopeLimiter") or "donor cell") if self.usePLM else 0
        if self.usePLM and self.fluxLimiter != 0:
            # gridsolver.lua:119: "are you sure you want to use flux and slope limiters at the same time?"
            raise ValueError("usePLM requires fluxLimiter='donor cell' (gridsolver.lua:119)")
        # gridsolver.lua:102-115: useCTU, switched off in 1-D
        self.useCTU = bool(args.get("useCTU", False)) and self.dim > 1
        if self.useCTU and self.usePLM != "plm cons":
            raise NotImplementedError("useCTU is built for usePLM='plm cons' (without PLM the reference's kernel rewrites UBuf itself, ctu.cl:77-84)")
        # boundary: cfg.boundary table wins, else what the initCond sets, else freeflow
        self.boundaryMethods = {}
        bargs = args.get("boundary") or {}
        for x in xNames:
            for mm in minmaxs:
                self.boundaryMethods[x + mm] = bargs.get(x + mm, "freeflow")
        if self.initCond.boundary and not bargs:
            self.setBoundaryMethods(self.initCond.boundary)

    def setBoundaryMethods(self, name):
        if isinstance(name, dict):
            self.boundaryMethods.update(name)
        else:
            for k in self.boundaryMethods:
                self.boundaryMethods[k] = name

    def boundaryIdList(self):
        out = []
        for x in xNames:
            for mm in minmaxs:
                m = self.boundaryMethods[x + mm]
                if isinstance(m, dict):     # {name = 'fixed', args = {...}} as in init/euler.lua:1859-1879
                    m = m["name"]
                if m not in boundaryIds:
                    raise NotImplementedError("boundary method %r is outside the hot-path scope" % (m,))
                out.append(boundaryIds[m])
        return out

    def fixedBoundaryStates(self):
        """-> {face index: list of numStates doubles} for the faces whose method is 'fixed' (gridsolver.lua:746-764).
        The reference takes a `fixedCode` generator; the states it writes in its own uses do not depend on the cell
        (init/euler.lua:1859-1879: consFromPrim of constants), so here args = {W = prims} or {U = conserved state}."""
        out = {}
        k = 0
        for x in xNames:
            for mm in minmaxs:
                m = self.boundaryMethods[x + mm]
                if isinstance(m, dict) and m["name"] == "fixed":
                    a = m.get("args") or {}
                    if "U" in a:
                        U = [float(v) for v in a["U"]]
                    elif "W" in a:
                        W = {key: np.float64(v) for key, v in a["W"].items()}
                        U = [float(v) for v in np.ravel(self.eqn.consArray(W))]
                    else:
                        raise ValueError("boundary 'fixed' needs args = {W = ...} or {U = ...} (gridsolver.lua:753)")
                    out[k] = U
                elif m == "fixed":
                    raise ValueError("you didn't provide any fixedCode arg for boundary==fixed (gridsolver.lua:753)")
                k += 1
        return out

    def cellPositions(self, part=None):
        """coord.lua:1421-1432: x = (i + .5 - g)/N * (max - min) + min on used axes, midpoint otherwise.
        With a slab decomposition these are the positions of this rank's (ghost-inclusive) slab.
        ``part = (axis, lo, hi)`` restricts one axis to the index range [lo, hi) of the ghost-inclusive local grid."""
        g = self.numGhost
        axes = []
        for j in range(3):
            if j < self.dim:
                i = np.arange(self.localGridSize[j], dtype=np.float64) + float(self.localOffset[j])
                a = (i + .5 - g) / float(self.sizeWithoutBorder[j]) * (self.maxs[j] - self.mins[j]) + self.mins[j]
                if part is not None and part[0] == j:
                    a = a[part[1]:part[2]]
                axes.append(a)
            else:
                axes.append(np.array([.5 * (self.maxs[j] + self.mins[j])]))
        z, y, x = np.meshgrid(axes[2], axes[1], axes[0], indexing="ij")
        return x, y, z      # each [Sz, Sy, Sx]

    def applyInitCond(self):
        """The initial condition is pointwise in the cell position (eqn.lua:622-632: one applyInitCond work-item per cell), so it is
        evaluated in chunks of planes of the slowest axis: the numpy temporaries of prims + consFromPrim are ~0.9 KB per cell, which at
        512^3 would be > 100 GB at once."""
        ax = self.dim - 1
        n = self.localGridSize[ax]
        per = self.localNumCells // n
        step = max(1, int(4e6 // max(per, 1)))
        S = self.localGridSize
        U = np.empty((S[2], S[1], S[0], self.eqn.numStates), dtype=np.float64)
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            x, y, z = self.cellPositions((ax, lo, hi))
            W = self.initCond.prims(x, y, z, self)
            sl = [slice(None)] * 3
            sl[2 - ax] = slice(lo, hi)
            U[sl[0], sl[1], sl[2]] = self.eqn.consArray(W)      # [.., numStates]
        self.setState(U)

    def setState(self, U):
        U = np.ascontiguousarray(U, dtype=np.float64).reshape(self.localNumCells, self.eqn.numStates)
        self.backend.set_state(U)

    def getState(self):
        """UBuf as float64 [Sz, Sy, Sx, numStates] (AoS cons_t order, ghost cells included); this rank's slab."""
        U = self.backend.get_state()
        return U.reshape(self.localGridSize[2], self.localGridSize[1], self.localGridSize[0], self.eqn.numStates)

    # ---- gridsolver.lua:1410-1470: GridSolver:saveBuffer / :save -- FITS dumps in the reference's own layout (hydro/fits.py), and
    #      the way back, so that a dump written by the reference (saveOnExit / save) can be loaded and compared
    def saveBuffer(self, U, basefn):
        from .. import fits
        real = np.float32 if self.real_bytes == 4 else np.float64
        fits.write_image(basefn + ".fits", fits.state_to_image(U, real))
        return basefn + ".fits"

    def save(self, prefix=None):
        """Writes <prefix>_UBuf.fits (the buffer of this path; the reference's loop also dumps its scratch buffers)."""
        return self.saveBuffer(self.getState(), (prefix + "_" if prefix else "") + "UBuf")

    def loadBuffer(self, filename):
        from .. import fits
        return fits.image_to_state(fits.read_image(filename), self.localGridSize)

    def load(self, prefix=None):
        self.setState(self.loadBuffer((prefix + "_" if prefix else "") + "UBuf.fits"))

    def getGlobalInterior(self):
        """Interior of the whole grid, gathered from every rank's slab (tests of the decomposed path)."""
        Ui = self.interior()
        return self.comm.gatherInterior(Ui, self.dim) if self.comm is not None else Ui

    def interior(self, U=None):
        U = self.getState() if U is None else U
        g = self.numGhost
        sl = [slice(g, -g) if j < self.dim else slice(None) for j in (2, 1, 0)]
        return U[sl[0], sl[1], sl[2]]

    def calcExactError(self, numStates=None):
        """gridsolver.lua:1337-1366, including its loop bounds (imax = gridSize - 2*ghost - 1 on the ghost-inclusive
        size, i.e. the last two interior cells per axis are skipped) and the division by the full interior volume."""
        if self.comm is not None:
            # the reference's loop bounds are on the global grid and the sum is over all cells: gather first
            raise NotImplementedError("calcExactError on a slab-decomposed solver: gather with getGlobalInterior() and use a single-device solver")
        eqn = self.eqn
        numStates = numStates or eqn.numIntStates
        U = self.getState()
        g = self.numGhost
        x, y, z = self.cellPositions()

        def rng(j):
            if j >= self.dim:
                return slice(0, 1)
            return slice(g, self.gridSize[j] - 2 * g - 1 + 1)
        sx, sy, sz = rng(0), rng(1), rng(2)
        exact = self.initCond.exactSolution(self.t, x[sz, sy, sx], self)
        err = 0.
        for j in range(numStates):
            err += float(np.sum(np.abs(U[sz, sy, sx, j] - exact[j])))
        vol = 1
        for n in self.sizeWithoutBorder:
            vol *= n
        return err / (numStates * vol)
