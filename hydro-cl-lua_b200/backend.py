"""CudaBackend: the device backend of the host-side solver mirror, through the C ABI (include/hydrob200.h).

It serves the same backend interface the tests' CPU oracle serves (oracle/oracle.py OracleBackend), so
``FiniteVolumeSolver(cfg)`` runs on the B200 and ``FiniteVolumeSolver(dict(cfg, backend=OracleBackend))`` runs the
CPU restatement of the reference.  This module never imports anything from oracle/.
"""
import ctypes as C

import numpy as np

from . import _lib as hb


class Context:
    """hb_ctx: one CUDA device + stream (the reference's CLEnv, hydro/app.lua:891-929)."""

    def __init__(self, device=0, real_bytes=8):
        self.L = hb.lib()
        self.h = hb.P()
        hb.check(self.L.hb_ctx_create(device, real_bytes, C.byref(self.h)))
        self.device = device
        self.real_bytes = real_bytes

    def close(self):
        if self.h:
            self.L.hb_ctx_destroy(self.h)
            self.h = hb.P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        hb.check(self.L.hb_sync(self.h))

    def deviceName(self):
        buf = C.create_string_buffer(256)
        hb.check(self.L.hb_device_name(self.h, buf, 256))
        return buf.value.decode()

    def smCount(self):
        n = C.c_int()
        hb.check(self.L.hb_device_sm_count(self.h, C.byref(n)))
        return n.value

    def timerStart(self):
        hb.check(self.L.hb_timer_start(self.h))

    def timerStop(self):
        ms = C.c_float()
        hb.check(self.L.hb_timer_stop(self.h, C.byref(ms)))
        return ms.value


def desc_from_solver(solver, strict_fp=False, use_graph=True, local_n=None, stage_kernel=0):
    d = hb.hb_fv_desc()
    d.eqn = solver.eqn.eqnId
    d.dim = solver.dim
    for i in range(3):
        d.global_n[i] = solver.sizeWithoutBorder[i]
        d.n[i] = (local_n or solver.sizeWithoutBorder)[i]
        d.mins[i] = solver.mins[i]
        d.maxs[i] = solver.maxs[i]
    d.use_plm = solver.plmId
    d.slope_limiter = solver.slopeLimiter
    d.flux_limiter = solver.fluxLimiter
    d.flux = solver.flux.fluxId
    d.flux_param = getattr(solver.flux, 'fluxParam', 0)
    for i, b in enumerate(solver.boundaryIdList()):
        d.bc[i] = b
    d.rk_order = solver.rkOrder
    for i in range(16):
        d.alphas[i] = solver.alphas[i]
        d.betas[i] = solver.betas[i]
    d.cfl = solver.cfl
    d.fixed_dt = solver.fixedDT if solver.useFixedDT else 0.
    d.use_fixed_dt = 1 if solver.useFixedDT else 0
    for i, v in enumerate(solver.eqn.eqnParams()):
        d.eqn_params[i] = v
    d.strict_fp = 1 if strict_fp else 0
    d.use_graph = 1 if use_graph else 0
    d.stage_kernel = stage_kernel
    d.use_ctu = 1 if getattr(solver, 'useCTU', False) else 0
    return d


class CudaBackend:
    """hb_fv: the fused finite-volume path on one GPU (or one slab of a decomposed grid)."""
    strict_fp = False
    use_graph = True

    def __init__(self, solver, device=0, comm=None):
        self.solver = solver
        self.L = hb.lib()
        args = solver.args
        self.ctx = args.get("ctx") or Context(device, solver.real_bytes)
        strict = args.get("strict_fp", self.strict_fp)
        graph = args.get("use_graph", self.use_graph)
        self.comm = comm
        self.desc = desc_from_solver(solver, strict, graph, solver.localSizeWithoutBorder, args.get("stage_kernel", 0))
        self.h = hb.P()
        src = args.get("eqnSource")
        if src:
            # the codegen seam: the equation's device functions as source, compiled by NVRTC into the marching kernels (hb_fv_create_from_source).
            # eqnSource = {name = include name, src = header text, type = class template name}; the host plug-in (args.eqn) supplies sizes / init
            log = C.create_string_buffer(1 << 16)
            rc = self.L.hb_fv_create_from_source(self.ctx.h, C.byref(self.desc), src["name"].encode(), src["src"].encode(), src["type"].encode(),
                                                 C.byref(self.h), log, len(log))
            self.compile_log = log.value.decode()
            if rc:
                raise hb.HydroB200Error(rc, self.L.hb_last_error().decode() + "\n" + self.compile_log)
        else:
            hb.check(self.L.hb_fv_create(self.ctx.h, C.byref(self.desc), C.byref(self.h)))
        ns, ni, nw = C.c_int(), C.c_int(), C.c_int()
        hb.check(self.L.hb_fv_num_states(self.h, C.byref(ns), C.byref(ni), C.byref(nw)))
        self.nS, self.nI, self.nW = ns.value, ni.value, nw.value
        self.ncells = self.L.hb_fv_num_cells(self.h)
        if comm is not None:
            hb.check(self.L.hb_fv_comm_init(self.h, comm.nranks, comm.rank, comm.uniqueId()))

    def close(self):
        if getattr(self, "h", None):
            self.L.hb_fv_destroy(self.h)
            self.h = hb.P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- backend interface (same as oracle.OracleBackend)
    def set_state(self, U):
        U = np.ascontiguousarray(U, dtype=np.float64)
        assert U.size == self.ncells * self.nS, (U.size, self.ncells, self.nS)
        hb.check(self.L.hb_fv_set_state(self.h, U.ctypes.data))
        self.ctx.sync()

    def get_state(self):
        U = np.empty((self.ncells, self.nS), dtype=np.float64)
        hb.check(self.L.hb_fv_get_state(self.h, U.ctypes.data))
        return U

    def boundary(self):
        hb.check(self.L.hb_fv_boundary(self.h))

    def add_op(self, kind, max_iters, stop_on_epsilon, stop_epsilon, param):
        o = hb.hb_op_desc(kind, max_iters, 1 if stop_on_epsilon else 0, stop_epsilon, param)
        idx = C.c_int()
        hb.check(self.L.hb_fv_add_op(self.h, C.byref(o), C.byref(idx)))
        return idx.value

    def ops_reset(self):
        hb.check(self.L.hb_fv_ops_reset(self.h))

    def op_info(self, op):
        it, res = C.c_int(), C.c_double()
        hb.check(self.L.hb_fv_op_info(self.h, int(op), C.byref(it), C.byref(res)))
        return it.value, res.value

    def set_fixed_boundary(self, face, U):
        a = (C.c_double * len(U))(*U)
        hb.check(self.L.hb_fv_set_fixed_boundary(self.h, int(face), a, len(U)))

    def constrainU(self):
        hb.check(self.L.hb_fv_constrainU(self.h))

    def init_derivs(self):
        hb.check(self.L.hb_fv_init_derivs(self.h))

    def calc_dt(self):
        dt = C.c_double()
        hb.check(self.L.hb_fv_calc_dt(self.h, C.byref(dt)))
        return dt.value

    def step(self, dt):
        hb.check(self.L.hb_fv_step(self.h, dt))

    def update(self, nsteps=1):
        hb.check(self.L.hb_fv_update(self.h, nsteps))
        return self.get_time()

    def get_time(self):
        t, dt = C.c_double(), C.c_double()
        hb.check(self.L.hb_fv_get_time(self.h, C.byref(t), C.byref(dt)))
        return t.value, dt.value

    def set_t(self, t):
        hb.check(self.L.hb_fv_set_time(self.h, t))

    def calc_deriv(self, dt):
        D = np.empty((self.ncells, self.nS), dtype=np.float64)
        hb.check(self.L.hb_fv_calc_deriv(self.h, dt, D.ctypes.data))
        return D

    # ---- extras
    def launch_count(self):
        n = C.c_longlong()
        hb.check(self.L.hb_fv_launch_count(self.h, C.byref(n)))
        return n.value

    def describe(self):
        buf = C.create_string_buffer(4096)
        hb.check(self.L.hb_fv_describe(self.h, buf, 4096))
        return buf.value.decode()


class CudaBackendStrict(CudaBackend):
    """Kernels built with -fmad=false: no FMA contraction, bit-comparable with the non-contracted oracle."""
    strict_fp = True
