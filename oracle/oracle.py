"""ctypes loader for the CPU parity oracle (oracle/hydro_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (hydro-cl-lua_b200/) never imports this module.

``OracleBackend`` implements the same backend interface the host-side solver mirror uses
(set_state/get_state/boundary/constrainU/calc_dt/step/update), so a parity test reads:

    ref = FiniteVolumeSolver(dict(cfg, backend=OracleBackend))     # CPU restatement of the reference
    gpu = FiniteVolumeSolver(cfg)                                  # CUDA product through the C-ABI
"""
import ctypes as C
import os
import subprocess

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))


class ho_desc(C.Structure):
    _fields_ = [
        ("eqn", C.c_int), ("dim", C.c_int), ("n", C.c_int * 3), ("real_bytes", C.c_int),
        ("use_plm", C.c_int), ("slope_limiter", C.c_int), ("flux_limiter", C.c_int),
        ("bc", C.c_int * 6), ("rk_order", C.c_int),
        ("alphas", C.c_double * 16), ("betas", C.c_double * 16),
        ("mins", C.c_double * 3), ("maxs", C.c_double * 3),
        ("cfl", C.c_double), ("fixed_dt", C.c_double), ("use_fixed_dt", C.c_int),
        ("gamma", C.c_double), ("rhoMin", C.c_double), ("PMin", C.c_double), ("mu0_eff", C.c_double),
        ("nthreads", C.c_int), ("global_n", C.c_int * 3), ("eqn_params", C.c_double * 16), ("flux", C.c_int), ("flux_param", C.c_int),
    ]


def build(force=False):
    """Compile the oracle with oracle/Makefile (g++, a few seconds)."""
    so = os.path.join(_here, "libhydro_oracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_here, "hydro_oracle.cpp")):
        subprocess.check_call(["make", "-C", _here, "-j2"], stdout=subprocess.DEVNULL)
    return so


_libs = {}


def _cpu_id():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def build_native():
    """The CPU-baseline build: g++ -O3 -march=native -ffp-contract=fast -fopenmp of the same restatement, compiled ON THE MACHINE THAT
    RUNS IT (oracle/_native/, git-ignored; rebuilt when the CPU model differs from the one it was built on).  Used by bench.py's
    cpu_baseline / --impl reference legs only: the parity tests use the two portable builds of oracle/Makefile."""
    d = os.path.join(_here, "_native")
    so, stamp = os.path.join(d, "libhydro_oracle_native.so"), os.path.join(d, "built_on.txt")
    src = os.path.join(_here, "hydro_oracle.cpp")
    ident = _cpu_id()
    fresh = os.path.exists(so) and os.path.exists(stamp) and open(stamp).read() == ident and os.path.getmtime(so) >= os.path.getmtime(src)
    if not fresh:
        os.makedirs(d, exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O3", "-march=native", "-ffp-contract=fast", "-fopenmp", "-shared", "-fPIC", "-w",
                               "-o", so, src])
        with open(stamp, "w") as f:
            f.write(ident)
    return so


def lib(fma=False):
    key = fma if fma == "native" else bool(fma)
    if key not in _libs:
        build()
        L = C.CDLL(build_native() if key == "native" else os.path.join(_here, "libhydro_oracle_fma.so" if fma else "libhydro_oracle.so"))
        L.ho_create.restype = C.c_void_p
        L.ho_create.argtypes = [C.POINTER(ho_desc)]
        for name in ("ho_destroy", "ho_boundary", "ho_constrainU", "ho_init_derivs"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.ho_num_states.argtypes = [C.c_void_p]
        L.ho_num_cells.argtypes = [C.c_void_p]
        L.ho_num_cells.restype = C.c_long
        L.ho_set_state.argtypes = [C.c_void_p, C.c_void_p]
        L.ho_set_fixed_boundary.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int]
        L.ho_set_fixed_boundary.restype = None
        L.ho_add_op.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        L.ho_ops_reset.argtypes = [C.c_void_p]
        L.ho_set_ctu.argtypes = [C.c_void_p, C.c_int]
        L.ho_ops_reset.restype = None
        L.ho_op_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.ho_op_info.restype = None
        L.ho_get_state.argtypes = [C.c_void_p, C.c_void_p]
        L.ho_calc_dt.argtypes = [C.c_void_p]
        L.ho_calc_dt.restype = C.c_double
        L.ho_update.argtypes = [C.c_void_p, C.c_int]
        L.ho_step.argtypes = [C.c_void_p, C.c_double]
        L.ho_get_t.argtypes = [C.c_void_p]
        L.ho_get_t.restype = C.c_double
        L.ho_get_dt.argtypes = [C.c_void_p]
        L.ho_get_dt.restype = C.c_double
        L.ho_set_t.argtypes = [C.c_void_p, C.c_double]
        L.ho_calc_deriv.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
        L.ho_roe_flux_test.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.ho_source_test.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ho_limiter.argtypes = [C.c_int, C.c_double]
        L.ho_limiter.restype = C.c_double
        L.ho_max_threads.restype = C.c_int
        _libs[key] = L
    return _libs[key]


def desc_from_solver(solver, nthreads=0):
    d = ho_desc()
    d.eqn = solver.eqn.eqnId
    d.dim = solver.dim
    for i in range(3):
        d.n[i] = solver.localSizeWithoutBorder[i]
        d.global_n[i] = solver.sizeWithoutBorder[i]
        d.mins[i] = solver.mins[i]
        d.maxs[i] = solver.maxs[i]
    d.real_bytes = solver.real_bytes
    d.use_plm = solver.plmId
    d.slope_limiter = solver.slopeLimiter
    d.flux_limiter = solver.fluxLimiter
    d.flux = solver.flux.fluxId
    d.flux_param = getattr(solver.flux, 'fluxParam', 0)
    bcs = solver.boundaryIdList()
    if getattr(solver, "comm", None) is not None:
        bcs = solver.comm.localBoundaryIds(bcs, solver.dim)     # faces owned by a neighbouring slab: 'none'
    for i, b in enumerate(bcs):
        d.bc[i] = b
    d.rk_order = solver.rkOrder
    for i in range(16):
        d.alphas[i] = solver.alphas[i]
        d.betas[i] = solver.betas[i]
    d.cfl = solver.cfl
    d.fixed_dt = solver.fixedDT if solver.useFixedDT else 0.
    d.use_fixed_dt = 1 if solver.useFixedDT else 0
    v = solver.eqn.vars
    for i, p in enumerate(solver.eqn.eqnParams()):
        d.eqn_params[i] = p
    d.gamma = v.get("heatCapacityRatio", 0.)
    d.rhoMin = v.get("rhoMin", 1e-7)
    d.PMin = v.get("PMin", 1e-7)
    d.mu0_eff = getattr(solver.eqn, "mu0_eff", 1.)
    d.nthreads = nthreads
    return d


class OracleBackend:
    """Backend interface of hydro/solver/solverbase.py, served by the CPU oracle."""
    fma = False
    nthreads = 0

    def __init__(self, solver):
        self.solver = solver
        self.L = lib(self.fma)
        self.desc = desc_from_solver(solver, self.nthreads)
        self.h = self.L.ho_create(C.byref(self.desc))
        if not self.h:
            raise RuntimeError("oracle: unsupported configuration")
        self.nS = self.L.ho_num_states(self.h)
        self.ncells = self.L.ho_num_cells(self.h)
        if getattr(solver, "useCTU", False) and self.L.ho_set_ctu(self.h, 1) != 0:
            raise RuntimeError("oracle: useCTU needs a PLM solver in more than one dimension")

    def __del__(self):
        try:
            if self.h:
                self.L.ho_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_state(self, U):
        U = np.ascontiguousarray(U, dtype=np.float64)
        assert U.size == self.ncells * self.nS
        self.L.ho_set_state(self.h, U.ctypes.data)

    def get_state(self):
        U = np.empty((self.ncells, self.nS), dtype=np.float64)
        self.L.ho_get_state(self.h, U.ctypes.data)
        return U

    def boundary(self):
        self.L.ho_boundary(self.h)

    def add_op(self, kind, max_iters, stop_on_epsilon, stop_epsilon, param):
        r = self.L.ho_add_op(self.h, int(kind), int(max_iters), 1 if stop_on_epsilon else 0, float(stop_epsilon), float(param))
        if r < 0:
            raise RuntimeError("oracle: op not available for this equation")
        return r

    def ops_reset(self):
        self.L.ho_ops_reset(self.h)

    def op_info(self, op):
        it, res = C.c_int(), C.c_double()
        self.L.ho_op_info(self.h, int(op), C.byref(it), C.byref(res))
        return it.value, res.value

    def set_fixed_boundary(self, face, U):
        a = (C.c_double * len(U))(*U)
        self.L.ho_set_fixed_boundary(self.h, int(face), a, len(U))

    def constrainU(self):
        self.L.ho_constrainU(self.h)

    def init_derivs(self):
        self.L.ho_init_derivs(self.h)

    def source_test(self, U):
        U = np.ascontiguousarray(U, dtype=np.float64)
        D = np.zeros(self.nS)
        self.L.ho_source_test(self.h, U.ctypes.data, D.ctypes.data)
        return D

    def calc_dt(self):
        return self.L.ho_calc_dt(self.h)

    def step(self, dt):
        self.L.ho_step(self.h, dt)

    def update(self, nsteps=1):
        self.L.ho_update(self.h, nsteps)
        return self.L.ho_get_t(self.h), self.L.ho_get_dt(self.h)

    def set_t(self, t):
        self.L.ho_set_t(self.h, t)

    def calc_deriv(self, dt):
        D = np.empty((self.ncells, self.nS), dtype=np.float64)
        self.L.ho_calc_deriv(self.h, D.ctypes.data, dt)
        return D

    def roe_flux_test(self, UL, UR, side):
        nI, nW = self.solver.eqn.numIntStates, self.solver.eqn.numWaves
        UL = np.ascontiguousarray(UL, dtype=np.float64)
        UR = np.ascontiguousarray(UR, dtype=np.float64)
        flux = np.zeros(self.nS)
        lam = np.zeros(nW)
        Lm = np.zeros((nW, nI))
        Rm = np.zeros((nI, nW))
        self.L.ho_roe_flux_test(self.h, UL.ctypes.data, UR.ctypes.data, side, flux.ctypes.data, lam.ctypes.data,
                                Lm.ctypes.data, Rm.ctypes.data)
        return flux, lam, Lm, Rm


class OracleBackendFMA(OracleBackend):
    """Same restatement compiled with -ffp-contract=fast (OpenCL's FP_CONTRACT default is ON)."""
    fma = True


def OracleBackendThreads(n, native=False):
    """`native`: the -O3 -march=native -ffp-contract=fast build (bench.py's CPU legs)"""
    return type("OracleBackendT%d" % n, (OracleBackend,), {"nthreads": n, "fma": "native" if native else False})
