"""GPU unit tests of the per-interface / per-cell device functions (hb_debug_eval) against the same functions
compiled for the host (tests/host_check), which tests/test_host_functions.py pins to the oracle."""
import numpy as np
import pytest

from test_host_functions import hc, make_oracle, random_states  # noqa: F401

pytestmark = pytest.mark.gpu


def dev_eval(eqnId, rb, strict, kind, side, n, params, aux, inp, out_per_item):
    from importlib import import_module
    backend = import_module("hydro-cl-lua_b200.backend")
    hb = import_module("hydro-cl-lua_b200._lib")
    ctx = backend.Context(0, rb)
    inp = np.ascontiguousarray(inp, dtype=np.float64)
    out = np.zeros(n * out_per_item)
    aux = np.ascontiguousarray(list(aux) + [0.] * (4 - len(aux)), dtype=np.float64)
    hb.check(ctx.L.hb_debug_eval(ctx.h, eqnId, 1 if strict else 0, kind, side, n, params.ctypes.data, aux.ctypes.data,
                                 inp.ctypes.data, inp.size, out.ctypes.data, out.size))
    return out.reshape(n, out_per_item)


def same(a, b):
    return (a == b) | (np.isnan(a) & np.isnan(b))


@pytest.mark.parametrize("eqn", ["euler", "mhd"])
@pytest.mark.parametrize("precision", ["double", "float"])
def test_device_functions_strict_match_host(hydrob200, oracle, hc, eqn, precision):
    S = make_oracle(hydrob200, oracle, eqn, precision)
    rb = 8 if precision == "double" else 4
    eid = S.eqn.eqnId
    nI = S.eqn.numIntStates
    params = np.array(S.eqn.eqnParams() + [0.] * 8, dtype=np.float64)
    rng = np.random.default_rng(4321)
    n = 256

    def rnd(a):
        return a.astype(np.float32).astype(np.float64) if precision == "float" else a

    U = random_states(eqn, 2 * n, rng)[:, :nI]
    U[::9, 0] = 1e-9
    U[::11, 4] = 1e-12
    U = np.ascontiguousarray(rnd(U))
    pairs = U.reshape(n, 2 * nI)
    report = []
    for side in range(3):
        got = dev_eval(eid, rb, True, 0, side, n, params, [], pairs, nI)
        ref = np.zeros((n, nI))
        for a in range(n):
            UL = np.ascontiguousarray(pairs[a, :nI]); UR = np.ascontiguousarray(pairs[a, nI:]); f = np.zeros(nI)
            hc.hc_roe_flux(eid, rb, side, params.ctypes.data, UL.ctypes.data, UR.ctypes.data, f.ctypes.data)
            ref[a] = f
        ok = same(got, ref)
        report.append(("roe side %d" % side, int((~ok).sum()), np.argwhere(~ok)[:4].tolist()))
    got = dev_eval(eid, rb, True, 1, 0, 2 * n, params, [], U, nI)
    ref = U.copy()
    for c in ref:
        u = np.ascontiguousarray(c); hc.hc_constrainU(eid, rb, params.ctypes.data, u.ctypes.data); c[:] = u
    ok = same(got, ref)
    report.append(("constrainU", int((~ok).sum()), np.argwhere(~ok)[:4].tolist()))
    dx = rnd(np.array([.01, .02, .03]))
    got = dev_eval(eid, rb, True, 2, 0, 2 * n, params, list(dx) + [3.], U, 1)[:, 0]
    ref = np.array([hc.hc_calc_dt_cell(eid, rb, params.ctypes.data, np.ascontiguousarray(c).ctypes.data, dx.ctypes.data, 3) for c in U])
    ok = same(got, ref)
    report.append(("calcDT", int((~ok).sum()), np.argwhere(~ok)[:4].tolist()))
    trip = rnd(rng.uniform(-1, 1, (n, 3)))
    for lim in (8, 18, 16):
        got = dev_eval(eid, rb, True, 3, 0, n, params, [float(lim)], trip, 1)[:, 0]
        ref = np.array([hc.hc_plm_half_slope(rb, lim, *t) for t in trip])
        report.append(("plm %d" % lim, int((~same(got, ref)).sum()), []))
    quads = np.ascontiguousarray(rnd(random_states(eqn, 4 * n, rng)[:, :nI]))
    dt_dx = float(rnd(np.array([.37]))[0])
    for side in range(3):
        got = dev_eval(eid, rb, True, 4, side, n, params, [18., dt_dx], quads.reshape(n, 4 * nI), nI)
        ref = np.zeros((n, nI))
        for a in range(n):
            q4 = np.ascontiguousarray(quads[4 * a:4 * a + 4]); f = np.zeros(nI)
            hc.hc_roe_flux_limited(eid, rb, side, params.ctypes.data, 18, dt_dx, q4.ctypes.data, f.ctypes.data)
            ref[a] = f
        ok = same(got, ref)
        report.append(("roe limited side %d" % side, int((~ok).sum()), np.argwhere(~ok)[:4].tolist()))
    bad = [r for r in report if r[1]]
    assert not bad, bad


def test_async_state_transfer_equals_blocking(hydrob200):
    """hb_fv_set_state_async / hb_fv_get_state_async / hb_fv_wait_transfers: three problems streamed through one solver (upload of
    the next overlapping the update and download of the previous) give exactly what the blocking calls give."""
    import ctypes as C
    from importlib import import_module
    hb = import_module("hydro-cl-lua_b200._lib")
    from cases import CASES
    cfg, _ = CASES["C4_sphere_rk4"]
    S = hydrob200.FiniteVolumeSolver(cfg)
    B, L = S.backend, S.backend.L
    U0 = S.getState().copy()
    n = U0.size
    inputs = [U0, U0 * np.array([1.1, 1., 1., 1., 1.2, 1.]), U0 * np.array([.9, 1., 1., 1., 1.05, 1.])]
    ref = []
    for Ui in inputs:
        S.setState(Ui); S.update(2); ref.append(S.getState().copy())
    hin = [C.c_void_p() for _ in inputs]; hout = [C.c_void_p() for _ in inputs]
    for h in hin + hout:
        hb.check(L.hb_host_alloc(n * 8, C.byref(h)))
    try:
        for h, Ui in zip(hin, inputs):
            C.memmove(h, np.ascontiguousarray(Ui).ctypes.data, n * 8)
        for hi, ho in zip(hin, hout):
            hb.check(L.hb_fv_set_state_async(B.h, hi))
            hb.check(L.hb_fv_update(B.h, 2))
            hb.check(L.hb_fv_get_state_async(B.h, ho))
        hb.check(L.hb_fv_wait_transfers(B.h))
        for ho, r in zip(hout, ref):
            got = np.frombuffer((C.c_double * n).from_address(ho.value), dtype=np.float64).reshape(r.shape)
            assert np.array_equal(got, r)
    finally:
        for h in hin + hout:
            hb.check(L.hb_host_free(h))
