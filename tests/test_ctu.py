"""useCTU (hydro/solver/ctu.cl, fvsolver.lua:246-272).  CPU: what the corner-transport correction is for -- an unsplit
PLM + forward-Euler update is unstable beyond CFL ~ .5 in 2-D, with the correction it runs at CFL .9 -- and conservation.
GPU: state parity of the unfused CTU kernel sequence against the oracle is in test_gpu_parity.py (cases F4_ctu_*); here the
launch structure and the refusal of unsupported combinations."""
import numpy as np
import pytest

KH = dict(eqn="euler", dim=2, gridSize=[48, 48], initCond="Kelvin-Helmholtz", usePLM="plm cons", slopeLimiter="minmod",
          integrator="forward Euler")


def total_variation(S):
    rho = S.getState()[0, 2:-2, 2:-2, 0]
    return np.abs(np.diff(rho, axis=1)).sum()


def test_ctu_is_stable_where_the_plain_update_is_not(hydrob200, oracle):
    A = hydrob200.FiniteVolumeSolver(dict(KH, cfl=.9, useCTU=True, backend=oracle.OracleBackend))
    B = hydrob200.FiniteVolumeSolver(dict(KH, cfl=.9, useCTU=False, backend=oracle.OracleBackend))
    for _ in range(60):
        A.update()
        B.update()
    # total variation of rho along x (0 initially: the shear layers are functions of y): the roll-up raises it to ~8; the plain
    # update's odd-even instability to hundreds before it produces NaNs
    assert np.isfinite(A.getState()).all() and total_variation(A) < 20.
    assert not np.isfinite(B.getState()).all() or total_variation(B) > 100.


def test_ctu_conserves_mass_and_changes_the_update(hydrob200, oracle):
    A = hydrob200.FiniteVolumeSolver(dict(KH, cfl=.3, useCTU=True, backend=oracle.OracleBackend))
    B = hydrob200.FiniteVolumeSolver(dict(KH, cfl=.3, backend=oracle.OracleBackend))
    m0 = A.getState()[0, 2:-2, 2:-2, 0].sum()
    for _ in range(10):
        A.update()
        B.update()
    assert abs(A.getState()[0, 2:-2, 2:-2, 0].sum() - m0) <= 1e-12 * m0        # periodic: flux form
    assert np.abs(A.getState() - B.getState()).max() > 1e-6                    # the correction is not a no-op


def test_ctu_is_off_in_1d_and_needs_plm(hydrob200, oracle):
    S = hydrob200.FiniteVolumeSolver(dict(eqn="euler", dim=1, gridSize=[32], initCond="Sod", usePLM="plm cons", useCTU=True,
                                          backend=oracle.OracleBackend))
    assert S.useCTU is False                                                   # gridsolver.lua:112-115
    with pytest.raises(NotImplementedError):
        hydrob200.FiniteVolumeSolver(dict(eqn="euler", dim=2, gridSize=[8, 8], useCTU=True, backend=oracle.OracleBackend))


@pytest.mark.gpu
def test_gpu_ctu_launch_structure_and_graph(hydrob200):
    cfg = dict(eqn="euler", dim=2, gridSize=[64, 48], initCond="Kelvin-Helmholtz", usePLM="plm cons", slopeLimiter="minmod",
               integrator="Runge-Kutta 2, TVD", cfl=.4, useCTU=True)
    A = hydrob200.FiniteVolumeSolver(dict(cfg, use_graph=True))
    B = hydrob200.FiniteVolumeSolver(dict(cfg, use_graph=False))
    assert "kernel=ctu(" in A.backend.describe()                                      # the unfused sequence, not a fused stage kernel
    n0 = B.backend.launch_count()
    A.update(4)
    for _ in range(4):
        B.update()
    assert np.array_equal(A.getState(), B.getState()) and A.t == B.t
    # per stage: calcLR, calcFlux, updateCTU, 2 * dim ghost fills of the face states, calcFlux, finish = 9 in 2-D (+ the state's ghost fill)
    assert (B.backend.launch_count() - n0) >= 4 * 2 * 10


@pytest.mark.gpu
def test_gpu_ctu_rejected_combinations(hydrob200):
    with pytest.raises(Exception):
        hydrob200.FiniteVolumeSolver(dict(eqn="euler", dim=2, gridSize=[16, 16], usePLM="plm cons", useCTU=True, stage_kernel=2))
