"""Shared parity cases: reduced-size versions of the BASELINE.json configs, sized so the CPU oracle
finishes each in seconds.  Grid sizes are deliberately not multiples of the CUDA tile shapes."""

CASES = {
    # C1: 1D Sod, Euler, Roe + superbee flux limiter, forward Euler, 256 cells (the reference's CPU-runnable config)
    "C1_sod_fe_superbee": (dict(eqn="euler", dim=1, gridSize=[256], initCond="Sod", fluxLimiter="superbee",
                                integrator="forward Euler", cfl=.3), 100),
    "C1_sod_fe_donor": (dict(eqn="euler", dim=1, gridSize=[300], initCond="Sod", fluxLimiter="donor cell",
                             integrator="forward Euler", cfl=.3), 50),
    "C1_sod_rk4_plm_mirror": (dict(eqn="euler", dim=1, gridSize=[200], initCond="Sod", usePLM="plm cons", slopeLimiter="minmod",
                                   integrator="Runge-Kutta 4", cfl=.3, boundary=dict(xmin="mirror", xmax="mirror")), 40),
    # C2: 2D Kelvin-Helmholtz, Euler, Roe + PLM, RK4-TVD, periodic
    "C2_kh_rk4tvd_minmod": (dict(eqn="euler", dim=2, gridSize=[80, 56], initCond="Kelvin-Helmholtz", usePLM="plm cons",
                                 slopeLimiter="minmod", integrator="Runge-Kutta 4, TVD", cfl=.15), 20),
    "C2_kh_rk4tvd_superbee": (dict(eqn="euler", dim=2, gridSize=[72, 40], initCond="Kelvin-Helmholtz", usePLM="plm cons",
                                   slopeLimiter="superbee", integrator="Runge-Kutta 4, TVD", cfl=.15), 20),
    "C2_sod2d_mirror_fluxlim": (dict(eqn="euler", dim=2, gridSize=[48, 36], initCond="Sod", fluxLimiter="superbee",
                                     integrator="Runge-Kutta 2, TVD", cfl=.15,
                                     boundary=dict(xmin="mirror", xmax="mirror", ymin="mirror", ymax="mirror")), 20),
    # C3: 2D Orszag-Tang, ideal MHD, Roe + PLM, RK3-TVD, periodic
    "C3_ot_rk3tvd": (dict(eqn="mhd", dim=2, gridSize=[72, 44], initCond="Orszag-Tang", usePLM="plm cons",
                          slopeLimiter="minmod", integrator="Runge-Kutta 3, TVD", cfl=.15), 20),
    "C3_briowu_1d_fluxlim": (dict(eqn="mhd", dim=1, gridSize=[256], initCond="Brio-Wu", fluxLimiter="superbee",
                                  integrator="forward Euler", cfl=.3), 50),
    # C4: 3D spherical blast, Euler, Roe + PLM, RK4, freeflow, domain +-2
    "C4_sphere_rk4": (dict(eqn="euler", dim=3, gridSize=[40, 20, 12], mins=[-2, -2, -2], maxs=[2, 2, 2], initCond="sphere",
                           usePLM="plm cons", slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.1), 10),
    "C4_sphere_rk4_mirror_periodic": (dict(eqn="euler", dim=3, gridSize=[24, 18, 12], mins=[-2, -2, -2], maxs=[2, 2, 2],
                                           initCond="sphere", usePLM="plm cons", slopeLimiter="minmod",
                                           integrator="Runge-Kutta 4", cfl=.1,
                                           boundary=dict(xmin="mirror", xmax="mirror", ymin="periodic", ymax="periodic",
                                                         zmin="freeflow", zmax="mirror")), 6),
    "C4_mhd3d_rk3": (dict(eqn="mhd", dim=3, gridSize=[20, 12, 10], initCond="Orszag-Tang", usePLM="plm cons",
                          slopeLimiter="minmod", integrator="Runge-Kutta 3, TVD", cfl=.1, mins=[-2, -2, -2], maxs=[2, 2, 2]), 5),
    # the volume > 1e-7 guard (fvsolver.cl:97): 512^3 on [-1,1]^3 would freeze; here a small grid on a tiny domain
    "frozen_volume_guard": (dict(eqn="euler", dim=3, gridSize=[12, 10, 8], mins=[-.01] * 3, maxs=[.01] * 3, initCond="sphere",
                                 usePLM="plm cons", slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.1), 2),
}

# gridsolver.lua:106: usePLM without a slopeLimiter key is 'donor cell' (zero slopes: first order), not minmod
CASES["C2_kh_plm_default_limiter"] = (dict(eqn="euler", dim=2, gridSize=[40, 24], initCond="Kelvin-Helmholtz", usePLM="plm cons",
                                           integrator="Runge-Kutta 2, TVD", cfl=.15), 8)

# the other interface fluxes of the calcFluxForInterface slot (hydro/flux/hll.cl 'Davis direct bounded', rusanov.cl)
CASES["F2_sod_hll_fe"] = (dict(eqn="euler", dim=1, gridSize=[256], initCond="Sod", flux="hll", integrator="forward Euler", cfl=.3), 60)
CASES["F2_kh_hll_plm_rk4"] = (dict(eqn="euler", dim=2, gridSize=[64, 40], initCond="Kelvin-Helmholtz", flux="hll", usePLM="plm cons",
                                   slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.15), 12)
CASES["F2_sphere_rusanov_3d"] = (dict(eqn="euler", dim=3, gridSize=[24, 18, 12], mins=[-2, -2, -2], maxs=[2, 2, 2], initCond="sphere",
                                      flux="rusanov", usePLM="plm cons", slopeLimiter="superbee", integrator="Runge-Kutta 3, TVD", cfl=.1,
                                      boundary=dict(xmin="mirror", xmax="mirror", ymin="periodic", ymax="periodic",
                                                    zmin="freeflow", zmax="freeflow")), 6)
CASES["F2_briowu_hll"] = (dict(eqn="mhd", dim=1, gridSize=[256], initCond="Brio-Wu", flux="hll", integrator="forward Euler", cfl=.3), 50)
CASES["F2_ot_rusanov_plm"] = (dict(eqn="mhd", dim=2, gridSize=[48, 36], initCond="Orszag-Tang", flux="rusanov", usePLM="plm cons",
                                   slopeLimiter="minmod", integrator="Runge-Kutta 3, TVD", cfl=.15), 12)
CASES["F2_ot_hll_3d"] = (dict(eqn="mhd", dim=3, gridSize=[16, 12, 10], initCond="Orszag-Tang", flux="hll", integrator="Runge-Kutta 2, TVD",
                              cfl=.1, mins=[-2, -2, -2], maxs=[2, 2, 2]), 5)

for _m in (0, 1, 2):
    CASES["F2_sod_hllc%d_fe" % _m] = (dict(eqn="euler", dim=1, gridSize=[256], initCond="Sod", flux="euler-hllc", hllcMethod=_m,
                                           integrator="forward Euler", cfl=.3), 60)
    CASES["F2_sphere_hllc%d_plm_3d" % _m] = (dict(eqn="euler", dim=3, gridSize=[20, 16, 12], mins=[-2, -2, -2], maxs=[2, 2, 2],
                                                  initCond="sphere", flux="euler-hllc", hllcMethod=_m, usePLM="plm cons",
                                                  slopeLimiter="minmod", integrator="Runge-Kutta 3, TVD", cfl=.1), 5)
CASES["F2_kh_hllc_2d"] = (dict(eqn="euler", dim=2, gridSize=[48, 40], initCond="Kelvin-Helmholtz", flux="euler-hllc",
                               integrator="Runge-Kutta 4", cfl=.15), 10)
# 'plm athena' (plm.cl:782-879), both face orders (as in the tree / as recorded), Euler
CASES["F1_sod_athena_fe"] = (dict(eqn="euler", dim=1, gridSize=[256], initCond="Sod", usePLM="plm athena", integrator="forward Euler", cfl=.3), 60)
CASES["F1_sod_athena_rec_rk4"] = (dict(eqn="euler", dim=1, gridSize=[200], initCond="Sod", usePLM="plm athena, recorded face order",
                                       integrator="Runge-Kutta 4", cfl=.3, boundary=dict(xmin="mirror", xmax="mirror")), 40)
CASES["F1_kh_athena_rec_rk4tvd"] = (dict(eqn="euler", dim=2, gridSize=[64, 40], initCond="Kelvin-Helmholtz",
                                         usePLM="plm athena, recorded face order", integrator="Runge-Kutta 4, TVD", cfl=.15), 12)
CASES["F1_sphere_athena_rec_3d"] = (dict(eqn="euler", dim=3, gridSize=[24, 18, 12], mins=[-2, -2, -2], maxs=[2, 2, 2], initCond="sphere",
                                         usePLM="plm athena, recorded face order", integrator="Runge-Kutta 3, TVD", cfl=.1), 6)
CASES["F1_sphere_athena_hll_3d"] = (dict(eqn="euler", dim=3, gridSize=[20, 14, 10], mins=[-2, -2, -2], maxs=[2, 2, 2], initCond="sphere",
                                         usePLM="plm athena", flux="hll", integrator="Runge-Kutta 2, TVD", cfl=.05), 4)

CASES["F1_briowu_athena_rk2"] = (dict(eqn="mhd", dim=1, gridSize=[256], initCond="Brio-Wu", usePLM="plm athena, recorded face order",
                                      integrator="Runge-Kutta 2, TVD", cfl=.3), 50)
CASES["F1_ot_athena_2d"] = (dict(eqn="mhd", dim=2, gridSize=[48, 36], initCond="Orszag-Tang", usePLM="plm athena, recorded face order",
                                 integrator="Runge-Kutta 3, TVD", cfl=.15), 10)
CASES["F1_ot_athena_tree_hll_2d"] = (dict(eqn="mhd", dim=2, gridSize=[40, 28], initCond="Orszag-Tang", usePLM="plm athena", flux="hll",
                                          integrator="Runge-Kutta 2, TVD", cfl=.1), 6)

# forward-Euler cases for the host-side (gloo) decomposition test; axis sizes divisible by 2 ranks
CASES["slab_fe_2d_periodic"] = (dict(eqn="euler", dim=2, gridSize=[24, 16], initCond="Kelvin-Helmholtz", usePLM="plm cons",
                                     slopeLimiter="minmod", integrator="forward Euler", cfl=.15), 6)
CASES["slab_fe_3d_mixed"] = (dict(eqn="mhd", dim=3, gridSize=[10, 8, 12], initCond="Orszag-Tang", fluxLimiter="superbee",
                                  integrator="forward Euler", cfl=.1, mins=[-2, -2, -2], maxs=[2, 2, 2],
                                  boundary=dict(xmin="periodic", xmax="periodic", ymin="mirror", ymax="mirror",
                                                zmin="freeflow", zmax="mirror")), 6)

# C5: 3D ADM Bona-Masso, Roe + superbee flux limiter (no PLM), RK4: gauge wave (periodic, domain +-.5) and warp bubble (freeflow)
_PER = dict(xmin="periodic", xmax="periodic", ymin="periodic", ymax="periodic", zmin="periodic", zmax="periodic")
ADM_CASES = {
    "C5_gauge_wave_rk4": (dict(eqn="adm3d", dim=3, gridSize=[20, 10, 9], mins=[-.5] * 3, maxs=[.5] * 3,
                               initCond="testbed - gauge wave", fluxLimiter="superbee", integrator="Runge-Kutta 4", cfl=.1,
                               boundary=_PER), 10),
    "C5_gauge_wave_f1": (dict(eqn="adm3d", dim=3, gridSize=[18, 7, 6], mins=[-.5] * 3, maxs=[.5] * 3,
                              initCond="testbed - gauge wave", fluxLimiter="superbee", integrator="Runge-Kutta 4", cfl=.1,
                              boundary=_PER, eqnArgs=dict(f_eqn="1")), 10),
    "C5_warp_bubble_rk4": (dict(eqn="adm3d", dim=3, gridSize=[14, 12, 10], initCond="Alcubierre warp bubble",
                                fluxLimiter="superbee", integrator="Runge-Kutta 4", cfl=.1), 10),
    "C5_gauge_wave_slab": (dict(eqn="adm3d", dim=3, gridSize=[12, 8, 8], mins=[-.5] * 3, maxs=[.5] * 3,
                                initCond="testbed - gauge wave", fluxLimiter="superbee", integrator="Runge-Kutta 4", cfl=.1,
                                boundary=_PER), 5),
    "C5_gauge_wave_donor_fe_2d": (dict(eqn="adm3d", dim=2, gridSize=[24, 10], mins=[-.5] * 3, maxs=[.5] * 3,
                                       initCond="testbed - gauge wave", fluxLimiter="donor cell", integrator="forward Euler",
                                       cfl=.1, boundary=dict(xmin="periodic", xmax="periodic", ymin="periodic", ymax="periodic")), 8),
}

FLOAT_CASES = ["C2_kh_rk4tvd_minmod", "C4_sphere_rk4", "C3_ot_rk3tvd", "C1_sod_fe_superbee", "F2_kh_hll_plm_rk4", "F2_ot_rusanov_plm"]

# slabs with >= 3 marching chunks per rank (KM = 64 planes in 3-D, 32 rows in 2-D): the overlapped exchange path of hb_fv.cu
CASES["slab_march3d_overlap"] = (dict(eqn="euler", dim=3, gridSize=[34, 10, 264], mins=[-2, -2, -2], maxs=[2, 2, 2], initCond="sphere",
                                      usePLM="plm cons", slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.1), 3)
CASES["slab_march3d_overlap_periodic"] = (dict(eqn="mhd", dim=3, gridSize=[33, 6, 264], mins=[-2, -2, -2], maxs=[2, 2, 2],
                                               initCond="Orszag-Tang", usePLM="plm cons", slopeLimiter="superbee",
                                               integrator="Runge-Kutta 3, TVD", cfl=.1,
                                               boundary=dict(xmin="periodic", xmax="periodic", ymin="mirror", ymax="mirror",
                                                             zmin="periodic", zmax="periodic")), 3)
CASES["slab_march2d_overlap"] = (dict(eqn="euler", dim=2, gridSize=[40, 140], initCond="Kelvin-Helmholtz", usePLM="plm cons",
                                      slopeLimiter="minmod", integrator="Runge-Kutta 4, TVD", cfl=.15), 4)

# thin slabs: 8 (5) planes per rank on 4 (8) ranks -- one marching chunk; the overlapped exchange splits it into rim and interior planes
CASES["slab_thin3d"] = (dict(eqn="euler", dim=3, gridSize=[34, 18, 32], mins=[-2, -2, -2], maxs=[2, 2, 2], initCond="sphere",
                             usePLM="plm cons", slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.1), 3)
CASES["slab_thin3d8"] = (dict(eqn="euler", dim=3, gridSize=[34, 18, 40], mins=[-2, -2, -2], maxs=[2, 2, 2], initCond="sphere",
                              usePLM="plm cons", slopeLimiter="minmod", integrator="Runge-Kutta 3, TVD", cfl=.1,
                              boundary=dict(xmin="mirror", xmax="mirror", ymin="freeflow", ymax="freeflow", zmin="periodic", zmax="periodic")), 3)
CASES["slab_thin2d_mhd"] = (dict(eqn="mhd", dim=2, gridSize=[70, 24], initCond="Orszag-Tang", usePLM="plm cons", slopeLimiter="minmod",
                                 integrator="Runge-Kutta 3, TVD", cfl=.15), 3)

# the remaining boundary methods of gridsolver.lua:746-846 (linear / quadratic extrapolation, fixed = Dirichlet state).
# These do not compose to a source-index map, so the GPU runs the reference's x, y, z passes (fill_ghosts_axis).
CASES["F4_sod_linear_1d"] = (dict(eqn="euler", dim=1, gridSize=[200], initCond="Sod", fluxLimiter="superbee", integrator="forward Euler", cfl=.3,
                                  boundary=dict(xmin="linear", xmax="quadratic")), 40)
CASES["F4_kh_linear_quadratic_2d"] = (dict(eqn="euler", dim=2, gridSize=[56, 40], initCond="Kelvin-Helmholtz", usePLM="plm cons", slopeLimiter="minmod",
                                           integrator="Runge-Kutta 4", cfl=.15,
                                           boundary=dict(xmin="linear", xmax="quadratic", ymin="quadratic", ymax="mirror")), 10)
CASES["F4_cavity_fixed_2d"] = (dict(eqn="euler", dim=2, gridSize=[40, 36], initCond="sphere", usePLM="plm cons", slopeLimiter="minmod",
                                    integrator="Runge-Kutta 3, TVD", cfl=.15,
                                    boundary=dict(xmin="mirror", xmax="mirror", ymin="mirror",
                                                  ymax=dict(name="fixed", args=dict(W=dict(rho=1., vx=2., vy=0., vz=0., P=1., ePot=0.))))), 12)
CASES["F4_sphere_mixed_3d"] = (dict(eqn="euler", dim=3, gridSize=[24, 18, 12], mins=[-2, -2, -2], maxs=[2, 2, 2], initCond="sphere",
                                    usePLM="plm cons", slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.1,
                                    boundary=dict(xmin="linear", xmax="freeflow", ymin="periodic", ymax="periodic", zmin="quadratic",
                                                  zmax=dict(name="fixed", args=dict(W=dict(rho=.01, vx=0., vy=0., vz=0., P=.01, ePot=0.))))), 5)
CASES["F4_ot_mhd_linear_2d"] = (dict(eqn="mhd", dim=2, gridSize=[48, 36], initCond="Orszag-Tang", usePLM="plm cons", slopeLimiter="minmod",
                                     integrator="Runge-Kutta 3, TVD", cfl=.15,
                                     boundary=dict(xmin="linear", xmax="linear", ymin="periodic", ymax="periodic")), 8)

# the ops either side of the step -- self-gravity (hydro/op/selfgrav.lua) inside every stage's addSource and NoDiv over the
# Jacobi relaxation (hydro/op/nodiv.lua, noDivPoissonSolver=jacobi) after the integrator.
CASES["F3_selfgrav_sphere_fe_2d"] = (dict(eqn="euler", dim=2, gridSize=[40, 28], initCond="sphere", fluxLimiter="superbee", integrator="forward Euler",
                                          cfl=.15, useGravity=True), 8)
CASES["F3_selfgrav_sphere_rk4_plm_3d"] = (dict(eqn="euler", dim=3, gridSize=[20, 14, 12], mins=[-2, -2, -2], maxs=[2, 2, 2], initCond="sphere",
                                               usePLM="plm cons", slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.1, useGravity=True,
                                               boundary=dict(xmin="mirror", xmax="mirror", ymin="freeflow", ymax="freeflow",
                                                             zmin="periodic", zmax="periodic")), 4)
CASES["F3_selfgrav_sod_rk2_1d"] = (dict(eqn="euler", dim=1, gridSize=[128], initCond="Sod", usePLM="plm cons", slopeLimiter="superbee",
                                        integrator="Runge-Kutta 2, TVD", cfl=.3, useGravity=True, opArgs=dict(maxIters=7, stopOnEpsilon=False)), 10)
CASES["F3_nodiv_ot_rk3_2d"] = (dict(eqn="mhd", dim=2, gridSize=[48, 36], initCond="Orszag-Tang", usePLM="plm cons", slopeLimiter="minmod",
                                    integrator="Runge-Kutta 3, TVD", cfl=.15, noDiv="jacobi"), 8)
CASES["F3_nodiv_selfgrav_ot_fe_3d"] = (dict(eqn="mhd", dim=3, gridSize=[16, 12, 10], initCond="Orszag-Tang", fluxLimiter="minmod",
                                            integrator="forward Euler", cfl=.1, mins=[-2, -2, -2], maxs=[2, 2, 2], noDiv="jacobi", useGravity=True), 4)

# useCTU (hydro/solver/ctu.cl, fvsolver.lua:246-272): PLM face states advanced half a step by the fluxes of all sides, boundary on
# the face states, second flux pass -- the reference's unfused kernel sequence on the GPU (hb_ctu_kernels.cuh)
CASES["F4_ctu_kh_fe_2d"] = (dict(eqn="euler", dim=2, gridSize=[48, 40], initCond="Kelvin-Helmholtz", usePLM="plm cons", slopeLimiter="minmod",
                                 integrator="forward Euler", cfl=.3, useCTU=True), 12)
CASES["F4_ctu_sphere_rk2_3d"] = (dict(eqn="euler", dim=3, gridSize=[20, 14, 12], mins=[-2, -2, -2], maxs=[2, 2, 2], initCond="sphere",
                                      usePLM="plm cons", slopeLimiter="minmod", integrator="Runge-Kutta 2, TVD", cfl=.2, useCTU=True,
                                      boundary=dict(xmin="mirror", xmax="freeflow", ymin="periodic", ymax="periodic",
                                                    zmin="freeflow", zmax="mirror")), 5)
CASES["F4_ctu_ot_mhd_fe_2d"] = (dict(eqn="mhd", dim=2, gridSize=[40, 32], initCond="Orszag-Tang", usePLM="plm cons", slopeLimiter="minmod",
                                     integrator="forward Euler", cfl=.3, useCTU=True), 8)
CASES["F4_ctu_hll_kh_rk4_2d"] = (dict(eqn="euler", dim=2, gridSize=[40, 28], initCond="Kelvin-Helmholtz", flux="hll", usePLM="plm cons",
                                      slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.3, useCTU=True), 6)

# more of plm.cl's reconstructions -- 'plm prim' (:191-253, slopes of the primitive variables) and 'piecewise constant' (:10-24)
CASES["F1_plm_prim_sod_rk2"] = (dict(eqn="euler", dim=1, gridSize=[200], initCond="Sod", usePLM="plm prim", slopeLimiter="superbee",
                                     integrator="Runge-Kutta 2, TVD", cfl=.3), 30)
CASES["F1_plm_prim_kh_rk4_2d"] = (dict(eqn="euler", dim=2, gridSize=[48, 36], initCond="Kelvin-Helmholtz", usePLM="plm prim", slopeLimiter="minmod",
                                       integrator="Runge-Kutta 4", cfl=.15), 10)
CASES["F1_plm_prim_sphere_hll_3d"] = (dict(eqn="euler", dim=3, gridSize=[20, 14, 12], mins=[-2, -2, -2], maxs=[2, 2, 2], initCond="sphere", flux="hll",
                                           usePLM="plm prim", slopeLimiter="minmod", integrator="Runge-Kutta 3, TVD", cfl=.1), 5)
CASES["F1_plm_prim_ot_mhd_2d"] = (dict(eqn="mhd", dim=2, gridSize=[40, 32], initCond="Orszag-Tang", usePLM="plm prim", slopeLimiter="minmod",
                                       integrator="Runge-Kutta 3, TVD", cfl=.15), 8)
CASES["F1_piecewise_constant_kh_2d"] = (dict(eqn="euler", dim=2, gridSize=[40, 28], initCond="Kelvin-Helmholtz", usePLM="piecewise constant",
                                             integrator="Runge-Kutta 2, TVD", cfl=.15), 10)
CASES["F3_selfgrav_kh_plm_rk4_2d"] = (dict(eqn="euler", dim=2, gridSize=[64, 40], initCond="Kelvin-Helmholtz", usePLM="plm cons", slopeLimiter="minmod",
                                           integrator="Runge-Kutta 4", cfl=.15, useGravity=True), 6)     # the 2-D marching kernel's gravity configuration
# 'plm cons with flux' (plm.cl:95-187, MUSCL-Hancock half step inside the reconstruction; the stage's dt enters the face states)
CASES["F1_plm_cons_flux_sod_fe"] = (dict(eqn="euler", dim=1, gridSize=[200], initCond="Sod", usePLM="plm cons with flux", slopeLimiter="minmod",
                                         integrator="forward Euler", cfl=.3), 40)
CASES["F1_plm_cons_flux_kh_rk2_2d"] = (dict(eqn="euler", dim=2, gridSize=[48, 36], initCond="Kelvin-Helmholtz", usePLM="plm cons with flux",
                                            slopeLimiter="minmod", integrator="Runge-Kutta 2, TVD", cfl=.15), 10)
CASES["F1_plm_cons_flux_ot_mhd_2d"] = (dict(eqn="mhd", dim=2, gridSize=[40, 32], initCond="Orszag-Tang", usePLM="plm cons with flux",
                                            slopeLimiter="minmod", integrator="forward Euler", cfl=.15), 8)
# the f1 remainder: 'plm eig' (plm.cl:256-427) and 'plm eig prim' / 'plm eig prim ref' (plm.cl:536-778), literal as the tree has them
# (", other face order": L and R exchanged -- the orientation in which the extrapolated states face the interface they are used at)
CASES["F1_plm_eig_sod_fe"] = (dict(eqn="euler", dim=1, gridSize=[200], initCond="Sod", usePLM="plm eig", slopeLimiter="minmod",
                                   integrator="forward Euler", cfl=.3), 40)
CASES["F1_plm_eig_kh_rk2_2d"] = (dict(eqn="euler", dim=2, gridSize=[48, 36], initCond="Kelvin-Helmholtz", usePLM="plm eig", slopeLimiter="superbee",
                                      integrator="Runge-Kutta 2, TVD", cfl=.15), 10)
CASES["F1_plm_eig_prim_sod_rk2"] = (dict(eqn="euler", dim=1, gridSize=[200], initCond="Sod", usePLM="plm eig prim",
                                         integrator="Runge-Kutta 2, TVD", cfl=.3), 30)
CASES["F1_plm_eig_prim_ref_kh_2d"] = (dict(eqn="euler", dim=2, gridSize=[48, 36], initCond="Kelvin-Helmholtz", usePLM="plm eig prim ref",
                                           integrator="forward Euler", cfl=.15), 10)
CASES["F1_plm_eig_prim_ref_other_sphere_3d"] = (dict(eqn="euler", dim=3, gridSize=[20, 14, 12], mins=[-2, -2, -2], maxs=[2, 2, 2], initCond="sphere",
                                                     usePLM="plm eig prim ref, other face order", integrator="Runge-Kutta 3, TVD", cfl=.1), 3)   # (the doubled central slope of plm.cl:640-647 drives the sphere's edge to negative pressure by step 5)
CASES["F1_plm_eig_prim_ref_sphere_3d"] = (dict(eqn="euler", dim=3, gridSize=[20, 14, 12], mins=[-2, -2, -2], maxs=[2, 2, 2], initCond="sphere",
                                               usePLM="plm eig prim ref", integrator="Runge-Kutta 3, TVD", cfl=.1), 5)
CASES["F1_plm_eig_prim_other_ot_mhd_2d"] = (dict(eqn="mhd", dim=2, gridSize=[40, 32], initCond="Orszag-Tang", usePLM="plm eig prim, other face order",
                                                 integrator="forward Euler", cfl=.15), 8)
CASES["F1_plm_eig_briowu_mhd"] = (dict(eqn="mhd", dim=1, gridSize=[200], initCond="Brio-Wu", usePLM="plm eig", slopeLimiter="minmod",
                                       integrator="forward Euler", cfl=.3), 30)
CASES["F3_selfgrav_linear_fixed_bc_2d"] = (dict(eqn="euler", dim=2, gridSize=[36, 28], initCond="sphere", usePLM="plm cons", slopeLimiter="minmod",
                                                integrator="Runge-Kutta 2, TVD", cfl=.15, useGravity=True,
                                                boundary=dict(xmin="linear", xmax="quadratic", ymin="mirror",
                                                              ymax=dict(name="fixed", args=dict(W=dict(rho=.01, vx=0., vy=0., vz=0., P=.01, ePot=0.))))), 6)

# float builds of the rows either side of the path (strict build bit-identical to the float oracle, production within 1e-5)
FLOAT_CASES += ["F3_nodiv_ot_rk3_2d", "F3_selfgrav_sphere_rk4_plm_3d", "F4_ctu_kh_fe_2d", "F4_kh_linear_quadratic_2d", "F1_plm_prim_kh_rk4_2d",
                "F1_plm_cons_flux_kh_rk2_2d"]

# ADM with the extrapolating boundary methods (the per-axis ghost passes on 51 variables)
ADM_CASES["C5_warp_bubble_linear_quadratic_bc"] = (dict(eqn="adm3d", dim=3, gridSize=[12, 10, 8], initCond="Alcubierre warp bubble", fluxLimiter="superbee",
                                                        integrator="Runge-Kutta 4", cfl=.1,
                                                        boundary=dict(xmin="linear", xmax="quadratic", ymin="freeflow", ymax="linear",
                                                                      zmin="quadratic", zmax="freeflow")), 6)
