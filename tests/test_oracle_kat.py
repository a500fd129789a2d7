"""Pins the CPU oracle against the reference's own known-answer values.

The reference's only recorded numbers for this path are the n=1024 L1 errors of rho (cfl .6, t>=.1) in the
trailing comments of tests/test-order/schemes.lua:61-126 (first column 'advect wave', second 'Sod'), produced
by tests/test-order/run.lua via GridSolver:calcExactError (gridsolver.lua:1337-1366).  The oracle reproduces
them through the same procedure: run update() until t >= duration, then the L1 error against the analytic
solution with the reference's loop bounds.

The rows recorded as usePLM='plm-cons' (schemes.lua:81-100) turn out to be the scheme the tree now calls
'plm cons with flux' (plm.cl:95-187): their Sod column is reproduced below to the recorded digits for every
limiter that is stable on the shock tube.  Rows recorded with 'plm-prim-alone' / 'plm-cons-alone' / 'plm-eig*'
are not reproduced by the current tree's variants and stay unpinned (the last tests of this file).
"""
import pytest

N = 1024
DURATION = .1


def run(hydrob200, oracle, ic, **kw):
    cfg = dict(eqn="euler", dim=1, gridSize=[N], initCond=ic, cfl=.6, backend=oracle.OracleBackend)
    cfg.update(kw)
    S = hydrob200.FiniteVolumeSolver(cfg)
    while S.t < DURATION:       # hydro/app.lua:1590: exit check before each update
        S.update()
    return S.calcExactError(1)


# (scheme kwargs, advect-wave KAT, Sod KAT, rel tol advect, rel tol Sod) -- schemes.lua line in the comment
KATS = [
    (dict(integrator="forward Euler", fluxLimiter="superbee"), 1.8741138235795e-06, 0.00029845606056438, 1e-9, 1e-12),      # :72
    (dict(integrator="forward Euler", fluxLimiter="donor cell"), 0.00029551600678436, 0.0025480145819915, 1e-11, 1e-12),    # :70
    (dict(integrator="forward Euler", fluxLimiter="minmod"), 2.549668825168e-06, 0.00084753894377874, 1e-9, 1e-12),        # :79
    (dict(integrator="forward Euler", fluxLimiter="monotized central"), 4.5584837524636e-07, 0.00045225285804202, 1e-8, 1e-11),  # :77
    (dict(integrator="forward Euler", fluxLimiter="van Leer"), 7.8024435458499e-07, 0.00054262570046035, 1e-8, 1e-11),     # :68
    (dict(integrator="forward Euler", fluxLimiter="Lax-Wendroff"), 7.5336775409067e-07, None, 1e-9, None),                  # :80
    (dict(integrator="forward Euler", usePLM="plm cons", slopeLimiter="donor cell"), 0.00029551600678436, 0.0025480145819915, 1e-11, 1e-12),  # :93
    (dict(integrator="Runge-Kutta 4", fluxLimiter="Lax-Wendroff"), 9.6682357219858e-05, None, 1e-10, None),                 # :110
    (dict(integrator="Runge-Kutta 4, TVD", fluxLimiter="Lax-Wendroff"), 9.6682357223262e-05, None, 1e-9, None),             # :109
    (dict(integrator="Runge-Kutta 3, TVD", fluxLimiter="Lax-Wendroff"), 9.6682375641956e-05, None, 1e-9, None),             # :107
    (dict(integrator="Runge-Kutta 2, TVD", fluxLimiter="Lax-Wendroff"), 9.6682087934274e-05, None, 1e-9, None),             # :112
    (dict(integrator="Runge-Kutta 4, non-TVD", fluxLimiter="Lax-Wendroff"), 9.6682357228525e-05, None, 1e-9, None),         # :108
    (dict(flux="hll", integrator="forward Euler"), 0.00037540541165354, 0.0029204942118918, 1e-11, 1e-12),                  # :35
    (dict(flux="euler-hllc", hllcMethod=0, integrator="forward Euler"), 0.00029551600678424, 0.0026034369564245, 1e-11, 1e-12),        # :39
    (dict(flux="euler-hllc", hllcMethod=1, integrator="forward Euler"), 0.0002955160067843, 0.0026034369564245, 1e-11, 1e-12),         # :47
    (dict(flux="euler-hllc", hllcMethod=2, integrator="forward Euler"), 0.00029551600678437, 0.0026034369564245, 1e-11, 1e-12),        # :55
    (dict(flux="euler-hllc", hllcMethod=0, integrator="Runge-Kutta 3, TVD"), 0.00039208124005629, 0.0031547573035925, 1e-10, 1e-12),   # :43
    (dict(flux="euler-hllc", hllcMethod=2, integrator="Runge-Kutta 3, TVD"), 0.00039208124005633, 0.0031547573035925, 1e-10, 1e-12),   # :59
    (dict(flux="euler-hllc", hllcMethod=1, integrator="Runge-Kutta 3, TVD"), 0.00039208124005635, 0.0031547573035925, 1e-10, 1e-12),   # :51
    # the other flux limiters of hydro/app.lua:614-635 (schemes.lua:61-78): all twenty limiter formulas are pinned
    (dict(integrator="forward Euler", fluxLimiter="smart"), 1.0911857176237e-06, 0.0006129313415327, 1e-8, 1e-12),   # :61
    (dict(integrator="forward Euler", fluxLimiter="ospre"), 0.00019705480525943, 0.0020319762206545, 1e-8, 1e-12),   # :62
    (dict(integrator="forward Euler", fluxLimiter="Fromm"), 1.5346799701183e-07, 0.00062670027933916, 1e-8, 1e-12),   # :63
    (dict(integrator="forward Euler", fluxLimiter="CHARM"), 1.4822461340873e-06, 0.00060082731534463, 1e-8, 1e-12),   # :64
    (dict(integrator="forward Euler", fluxLimiter="van Albada 1"), 1.2462029419466e-06, 0.00065025946311501, 1e-8, 1e-12),   # :65
    (dict(integrator="forward Euler", fluxLimiter="Barth-Jespersen"), 4.808764902909e-07, 0.00045706811880196, 1e-8, 1e-12),   # :66
    (dict(integrator="forward Euler", fluxLimiter="Beam-Warming"), 1.0595112307524e-06, 0.0013601751183648, 1e-8, 1e-12),   # :67
    (dict(integrator="forward Euler", fluxLimiter="van Albada 2"), 2.3724738671197e-06, 0.00081184935099592, 1e-8, 1e-12),   # :69
    (dict(integrator="forward Euler", fluxLimiter="Oshker"), 2.4949277440571e-06, 0.00067671846961815, 1e-8, 1e-12),   # :71
    (dict(integrator="forward Euler", fluxLimiter="Sweby"), 2.0705889490385e-06, 0.00041546702447137, 1e-8, 1e-12),   # :73
    (dict(integrator="forward Euler", fluxLimiter="HQUICK"), 1.3626347787182e-06, 0.00059526337064112, 1e-8, 1e-12),   # :74
    (dict(integrator="forward Euler", fluxLimiter="Koren"), 1.0448963734558e-06, 0.00050202680619077, 1e-8, 1e-12),   # :75
    (dict(integrator="forward Euler", fluxLimiter="HCUS"), 1.1792669058976e-06, 0.00055928895482656, 1e-8, 1e-12),   # :76
    (dict(integrator="forward Euler", fluxLimiter="UMIST"), 1.1751903573017e-06, 0.00061938621334895, 1e-8, 1e-12),   # :78
    # SURVEY 8f1 'plm athena'.  The tree assigns the face states the other way round (plm.cl:877-878: L = cons(Wrv), R = cons(Wlv))
    # than the version these rows were recorded with; with L = left, R = right every row is reproduced.
    (dict(usePLM="plm athena, recorded face order", integrator="forward Euler"), 9.7002822784791e-05, 0.00093771140713331, 1e-10, 1e-11),   # :97
    (dict(usePLM="plm athena, recorded face order", integrator="Runge-Kutta 4"), 1.2279744814311e-06, 0.00069783409481987, 1e-9, 1e-11),     # :116
    (dict(usePLM="plm athena, recorded face order", integrator="Runge-Kutta 2"), 1.3132898617302e-06, 0.00067817159415456, 1e-9, 1e-11),     # :112
    (dict(usePLM="plm athena, recorded face order", integrator="Runge-Kutta 3, TVD"), 1.226178781197e-06, 0.00070049425851808, 1e-9, 1e-11),  # :120
    (dict(usePLM="plm athena, recorded face order", integrator="Runge-Kutta 4, TVD"), 1.2279971850421e-06, 0.00069848808747002, 1e-9, 1e-11),  # :119
    # the remaining integrator rows (schemes.lua:104-125): every Butcher tableau of hydro/int/all.lua is pinned through them
    (dict(usePLM="plm athena, recorded face order", integrator="Runge-Kutta 2 Ralston"), 1.3128886603164e-06, 0.0007200643702763, 1e-9, 1e-11),       # :118
    (dict(usePLM="plm athena, recorded face order", integrator="Runge-Kutta 2, TVD"), 1.3118684200161e-06, 0.00072699746948317, 1e-9, 1e-11),         # :119
    (dict(usePLM="plm athena, recorded face order", integrator="Runge-Kutta 2 Heun"), 1.3118684201669e-06, 0.00072699746948318, 1e-9, 1e-11),         # :120
    (dict(usePLM="plm athena, recorded face order", integrator="Runge-Kutta 4, non-TVD"), 1.2279744811973e-06, 0.0006978340948199, 1e-8, 1e-11),      # :122
    (dict(usePLM="plm athena, recorded face order", integrator="Runge-Kutta 4, 3/8ths rule"), 7.2560410495091e-05, 0.00073737230229337, 1e-11, 1e-11),  # :125
    (dict(usePLM="plm athena, recorded face order", integrator="Runge-Kutta 3"), 0.042519919280547, 0.013615227412304, 1e-11, 1e-11),                 # :116
    (dict(integrator="Runge-Kutta 2 Heun", fluxLimiter="Lax-Wendroff"), 9.6682087934198e-05, None, 1e-10, None),            # :113
    (dict(integrator="Runge-Kutta 2 Ralston", fluxLimiter="Lax-Wendroff"), 9.6682087934305e-05, None, 1e-10, None),         # :114
    (dict(integrator="Runge-Kutta 2", fluxLimiter="Lax-Wendroff"), 9.6682087934377e-05, None, 1e-10, None),                 # :115
    (dict(integrator="Runge-Kutta 4, 3/8ths rule", fluxLimiter="Lax-Wendroff"), 2.4183176943556e-05, None, 1e-10, None),    # :116
    (dict(integrator="Runge-Kutta 3", fluxLimiter="Lax-Wendroff"), 0.042514748138836, None, 1e-11, None),                   # :106
    (dict(integrator="Runge-Kutta 2, non-TVD", fluxLimiter="Lax-Wendroff"), 9.6682087934282e-05, None, 1e-10, None),        # :111
]


@pytest.mark.parametrize("kw,kat_adv,kat_sod,tol_adv,tol_sod", KATS, ids=[str(sorted(str(v) for v in k[0].values())) for k in KATS])
def test_schemes_lua_kat(hydrob200, oracle, kw, kat_adv, kat_sod, tol_adv, tol_sod):
    got = run(hydrob200, oracle, "advect wave", **kw)
    assert abs(got - kat_adv) <= tol_adv * kat_adv, (got, kat_adv)
    if kat_sod is not None:
        got = run(hydrob200, oracle, "Sod", **kw)
        assert abs(got - kat_sod) <= tol_sod * kat_sod, (got, kat_sod)


# The rows the reference recorded as usePLM='plm-cons' with a slope limiter (schemes.lua:81-100; Roe, forward Euler) are the scheme its
# tree now calls 'plm cons with flux' (plm.cl:95-187, the MUSCL-Hancock variant): the oracle reproduces their Sod column for every
# limiter that is stable on the shock tube, to the digits recorded.  (The advect-wave column of these rows is reproduced only where the
# limited scheme is well conditioned on smooth data -- ospre, donor cell; with the compressive limiters the recorded L1 errors are
# O(1e-2) of an unstable run and depend on rounding.)  This pins plm.cl:95-187 and eleven more of the limiter formulas of
# hydro/app.lua:614-635 on the reference's own numbers.
PLM_FLUX_KATS = [
    # (slopeLimiter, advect-wave KAT or None, Sod KAT, rel tol advect, rel tol Sod)                       schemes.lua line
    ("minmod", None, 0.0013513761228327, None, 1e-12),                                                  # :100
    ("ospre", 0.00013263354368475, 0.0016309138954085, 1e-10, 1e-12),                                   # :96
    ("donor cell", 0.00029551600678436, 0.0025480145819915, 1e-11, 1e-12),                              # :93
    ("van Albada 1", None, 0.0014345138627451, None, 1e-11),                                            # :98
    ("UMIST", None, 0.0030517363635154, None, 1e-10),                                                   # :94
    ("Oshker", None, 0.002464024257206, None, 1e-9),                                                    # :97
    ("HQUICK", None, 0.0043568816756108, None, 1e-9),                                                   # :91
    ("Koren", None, 0.0045319629410316, None, 1e-8),                                                    # :90
    ("HCUS", None, 0.0034376341809745, None, 1e-8),                                                     # :92
    ("monotized central", None, 0.0048356399731734, None, 1e-8),                                        # :89
    ("superbee", None, 0.0054859189942998, None, 1e-6),                                                 # :88
    ("Sweby", None, 0.0025329368985951, None, 1e-6),                                                    # :95
]


@pytest.mark.parametrize("lim,kat_adv,kat_sod,tol_adv,tol_sod", PLM_FLUX_KATS, ids=[k[0] for k in PLM_FLUX_KATS])
def test_schemes_lua_plm_cons_rows_are_plm_cons_with_flux(hydrob200, oracle, lim, kat_adv, kat_sod, tol_adv, tol_sod):
    kw = dict(usePLM="plm cons with flux", slopeLimiter=lim, integrator="forward Euler")
    got = run(hydrob200, oracle, "Sod", **kw)
    assert abs(got - kat_sod) <= tol_sod * kat_sod, (got, kat_sod)
    if kat_adv is not None:
        got = run(hydrob200, oracle, "advect wave", **kw)
        assert abs(got - kat_adv) <= tol_adv * kat_adv, (got, kat_adv)


# ---- 'plm eig', 'plm eig prim', 'plm eig prim ref' (plm.cl:256-427, 536-778): literal restatements of the tree's code.  The rows the
# reference recorded under these names (schemes.lua:101-103: 2.9593e-4 / 1.4406e-3, 2.9593e-4 / 1.4316e-3, 2.9590e-4 / 1.2224e-3) come
# from an earlier version of that code and are NOT reproduced -- in either face order (the closest: 'plm eig prim ref' with L and R
# exchanged, 2.95928e-4 / 1.7355e-3).  What can be held against the reference's numbers are two reductions of the current code:
def test_plm_eig_with_zero_slopes_is_the_recorded_donor_cell_row(hydrob200, oracle):
    """'plm eig' limits the characteristic differences with the modular slope limiter; with 'donor cell' (phi = 0) every slope is zero
    and the scheme is first-order upwind plus a vanishing half-step flux difference: schemes.lua:93's numbers."""
    kw = dict(usePLM="plm eig", slopeLimiter="donor cell", integrator="forward Euler")
    got = run(hydrob200, oracle, "advect wave", **kw)
    assert abs(got - 0.00029551600678436) <= 1e-11 * 0.00029551600678436, got
    got = run(hydrob200, oracle, "Sod", **kw)
    assert abs(got - 0.0025480145819915) <= 1e-12 * 0.0025480145819915, got


def test_plm_eig_prim_tree_face_order_is_first_order_on_an_entropy_wave(hydrob200, oracle):
    """In the tree's assignment (plm.cl:703-704: result->L = the state extrapolated towards +side) the flux kernel reads, at every
    interface, the face states whose characteristic slopes were zeroed for the direction the wave travels in: a pure entropy wave
    moving in +x sees cell averages only, i.e. the recorded donor-cell error (schemes.lua:70) to rounding."""
    got = run(hydrob200, oracle, "advect wave", usePLM="plm eig prim", integrator="forward Euler")
    assert abs(got - 0.00029551600678436) <= 1e-9 * 0.00029551600678436, got


@pytest.mark.parametrize("plm,kat_adv,kat_sod", [("plm eig prim ref", 0.00029592694908829, 0.0014406203418454),
                                                 ("plm eig prim", 0.00029593080182222, 0.0014316307885825)])
def test_plm_eig_prim_recorded_rows_are_from_other_code(hydrob200, oracle, plm, kat_adv, kat_sod):
    """Recorded, not reproduced: kept as a test so that the statement in DESIGN.md section 7 stays checked."""
    for order in ("", ", other face order"):
        got = run(hydrob200, oracle, "Sod", usePLM=plm + order, integrator="forward Euler")
        assert abs(got - kat_sod) > 1e-2 * kat_sod, (plm + order, got, kat_sod)
