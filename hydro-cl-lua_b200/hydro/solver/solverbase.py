"""SolverBase: host orchestration of one solver (hydro/solver/solverbase.lua), device work delegated to a backend.

Reference: ctor args solverbase.lua:351-395,795-815 (integrator, fixedDT, cfl default .5, fluxLimiter),
``update`` :3026-3190, ``calcDT`` :3004-3023, ``step`` :3193-3238, ``constrainU`` :2116-2127,
``resetState`` :2094-2114 + hydro/init/init.lua:206-248.

Only the OpenCL-touching parts change: where the reference enqueues ~67 kernels per RK4 step through
lua-opencl, this class calls the C-ABI fused path (hb_fv_update / hb_fv_step, include/hydrob200.h)
through its backend.  The backend must be the CUDA library; there is no CPU fallback.
"""
import numpy as np

from .. import app as hydro_app
from ..eqn import eqns
from ..init.einstein import initConds as einsteinInitConds
from ..init.euler import initConds as eulerInitConds

initConds = dict(eulerInitConds)
initConds.update(einsteinInitConds)
from ..int import all as int_all


class SolverBase:
    name = "SolverBase"

    def __init__(self, args):
        args = dict(args)
        self.args = args
        self.app = args.get("app")
        self.dim = int(args.get("dim", 1))
        self.t = 0.
        self.dt = 0.
        self.ops = []                                   # parity contract: no ops
        self.initMeshVars(args)
        self.initObjs(args)
        self.backend = self.createBackend(args)
        for face, U in self.fixedBoundaryStates().items():
            self.backend.set_fixed_boundary(face, U)
        self.createOps(args)
        self.resetState()

    def fixedBoundaryStates(self):
        return {}

    def createOps(self, args):
        """solver.ops (euler.lua:179-188, mhd.lua:110-122).  The reference always inserts SelfGrav and switches it with
        solver.useGravity (selfgrav.lua:22,113), which an initial condition may set; NoDiv joins mhd in more than one dimension.
        Here: useGravity=True adds SelfGrav; noDiv='jacobi' adds NoDiv with the Jacobi parent (the default krylov parent is the
        reference's un-vendored 'solver' library: not built, so NoDiv stays off unless asked for)."""
        from ..op import NoDiv, SelfGrav
        opArgs = args.get("opArgs") or {}
        self.useGravity = bool(args.get("useGravity", False))
        name = getattr(self.eqn, "name", None)
        if args.get("noDiv"):
            if args["noDiv"] != "jacobi":
                raise NotImplementedError("noDiv=%r: only the Jacobi parent (noDivPoissonSolver=jacobi) is built" % (args["noDiv"],))
            if name != "mhd" or self.dim < 2:
                raise ValueError("NoDiv is an op of the mhd equation in more than one dimension (mhd.lua:113-119)")
            self.ops.append(NoDiv(self, **opArgs))
        if self.useGravity:
            if name not in ("euler", "mhd"):
                raise ValueError("selfgrav is an op of the euler and mhd equations")
            self.ops.append(SelfGrav(self, **opArgs))
        for op in self.ops:
            op.register(self.backend)

    # ---- solverbase.lua:511-556 / gridsolver.lua:60-96 (overridden by GridSolver)
    def initMeshVars(self, args):
        pass

    # ---- solverbase.lua:775-892
    def initObjs(self, args):
        self.integratorName = args.get("integrator", "forward Euler")
        self.rkOrder, self.alphas, self.betas = int_all.tableau(self.integratorName)
        self.useFixedDT = args.get("fixedDT") is not None
        self.fixedDT = args.get("fixedDT") if self.useFixedDT else .001
        self.cfl = args.get("cfl", .5)
        self.fluxLimiter = hydro_app.limiterIndex(args.get("fluxLimiter", "donor cell"))
        self.real_bytes = hydro_app.realBytes(args.get("precision", "double" if not args.get("float") else "float"))
        self.createEqn(args)

    def createEqn(self, args):
        name = args.get("eqn", "euler")
        if name not in eqns:
            raise NotImplementedError("eqn %r is outside the hot-path scope (have: %s)" % (name, ", ".join(eqns)))
        self.eqn = eqns[name](self, args.get("eqnArgs"))
        icname = args.get("initCond", "Sod")
        if icname not in initConds:
            raise NotImplementedError("initCond %r not provided (have: %s)" % (icname, ", ".join(initConds)))
        self.initCond = initConds[icname](args.get("initCondArgs"))
        self.eqn.applySolverVars(self.initCond.solverVars)       # solverbase.lua:1530-1536

    @property
    def heatCapacityRatio(self):
        return self.eqn.vars["heatCapacityRatio"]

    def createBackend(self, args):
        raise NotImplementedError

    # ---- hot loop ---------------------------------------------------------------------------
    def calcDT(self):
        """solverbase.lua:3004-3023: dt = cfl * min over interior cells of dx/|lambda|max, or fixedDT."""
        if self.useFixedDT:
            return self.fixedDT
        dt = self.backend.calc_dt()
        if not np.isfinite(dt):
            print("got a bad dt at time %r" % self.t)
        self.fixedDT = dt
        return dt

    def step(self, dt):
        """solverbase.lua:3193-3238: integrator:integrate(dt, calcDeriv [+ addSource]); ops: none."""
        self.backend.step(dt)

    def update(self, nsteps=1):
        """solverbase.lua:3026-3190.  ``nsteps`` > 1 runs whole updates back to back on the device with the
        dt reduction kept device-resident (no host read-back between steps)."""
        self.t, self.dt = self.backend.update(nsteps)

    def boundary(self):
        self.backend.boundary()

    def constrainU(self):
        """solverbase.lua:2116-2127: constrainU kernel then boundary()."""
        self.backend.constrainU()

    def resetState(self):
        """gridsolver.lua:519 -> solverbase.lua:2094 -> init.lua:206-248: applyInitCond, boundary, constrainU."""
        self.t = 0.
        self.backend.set_t(0.)
        self.applyInitCond()
        if getattr(self.eqn, "name", None) == "adm3d":
            # init.lua:231-235: equations with an initDerivs kernel (adm3d.cl:196-243): boundary, then finite-difference a_l, d_lll, V_l
            self.boundary()
            self.backend.init_derivs()
        self.boundary()
        if self.ops:
            # solverbase.lua:2106-2111: op:resetState() then boundary(), for every op
            self.backend.ops_reset()
        self.constrainU()

    def applyInitCond(self):
        raise NotImplementedError
