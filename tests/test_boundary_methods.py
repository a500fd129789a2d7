"""Boundary methods of hydro/solver/gridsolver.lua:638-846, checked three ways:
  * the oracle's ghost fill against an independent pure-numpy restatement of the reference's kernel loop
    (`for j < numGhost { min face; max face }`, boundary_x then _y then _z, gridsolver.lua:1100-1190, 1272-1314) -- CPU;
  * the analytic property: linear extrapolation reproduces a linear field, quadratic a quadratic one -- CPU;
  * the CUDA ghost fill (per-axis passes, fill_ghosts_axis) against the oracle, value for value -- GPU.
"""
import numpy as np
import pytest


def numpy_boundary(U, dim, bc, fixed, flips):
    """U[k, j, i, var] (ghost-inclusive); bc = 6 method names; fixed = {face: state}; flips(var, side) -> bool (mirror)."""
    g = 2
    U = U.copy()
    for side in range(dim):
        ax = 2 - side                      # numpy axis of this side
        A = np.moveaxis(U, ax, 0)          # view: A[index along side, ...]
        S = A.shape[0]
        N = S - 2 * g
        for j in range(g):
            for mm in range(2):
                m = bc[2 * side + mm]
                if m == "periodic":
                    if mm == 0:
                        A[j] = A[g + (j - g + 2 * N) % N]
                    else:
                        A[S - 1 - j] = A[g + (g - 1 - j) % N]
                elif m == "mirror":
                    dst, src = (j, 2 * g - 1 - j) if mm == 0 else (S - g + j, S - g - 1 - j)
                    A[dst] = A[src]
                    for q in range(U.shape[-1]):
                        if flips(q, side):
                            A[dst][..., q] = -1. * A[dst][..., q]
                elif m == "freeflow":
                    if mm == 0:
                        A[j] = A[g]
                    else:
                        A[S - g + j] = A[S - g - 1]
                elif m == "linear":
                    if mm == 0:
                        A[g - j - 1] = 2. * A[g - j] - A[g - j + 1]
                    else:
                        A[S - g + j] = 2. * A[S - g + j - 1] - A[S - g + j - 2]
                elif m == "quadratic":
                    if mm == 0:
                        A[g - j - 1] = 3. * A[g - j] - 3. * A[g - j + 1] + A[g - j + 2]
                    else:
                        A[S - g + j] = 3. * A[S - g + j - 1] - 3. * A[S - g + j - 2] + A[S - g + j - 3]
                elif m == "fixed":
                    A[j if mm == 0 else S - g + j] = np.asarray(fixed[2 * side + mm])
    return U


LID = dict(name="fixed", args=dict(W=dict(rho=1., vx=2., vy=0., vz=0., P=1., ePot=0.)))
CONFIGS = {
    "1d": (dict(eqn="euler", dim=1, gridSize=[17], boundary=dict(xmin="linear", xmax="quadratic")), ["linear", "quadratic"] + ["freeflow"] * 4),
    "2d": (dict(eqn="euler", dim=2, gridSize=[13, 9], boundary=dict(xmin="quadratic", xmax="mirror", ymin="linear", ymax=LID)),
           ["quadratic", "mirror", "linear", "fixed", "freeflow", "freeflow"]),
    "3d": (dict(eqn="euler", dim=3, gridSize=[9, 7, 6], boundary=dict(xmin="linear", xmax="periodic", ymin="periodic", ymax="quadratic", zmin=LID, zmax="linear")),
           ["linear", "periodic", "periodic", "quadratic", "fixed", "linear"]),
    "3d_mhd": (dict(eqn="mhd", dim=3, gridSize=[8, 7, 6], initCond="Orszag-Tang",
                    boundary=dict(xmin="mirror", xmax="linear", ymin="quadratic", ymax="mirror", zmin="freeflow", zmax="quadratic")),
               ["mirror", "linear", "quadratic", "mirror", "freeflow", "quadratic"]),
}


def _flips(eqn):
    if eqn == "mhd":
        return lambda q, side: q == 1 + side or q == 5 + side
    return lambda q, side: q == 1 + side


def _fixed_states(S):
    return S.fixedBoundaryStates()


def _random_state(S, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(.5, 2., size=(S.numCells, S.eqn.numStates))


@pytest.mark.parametrize("name", list(CONFIGS))
def test_oracle_boundary_equals_numpy_restatement(hydrob200, oracle, name):
    cfg, bc = CONFIGS[name]
    S = hydrob200.FiniteVolumeSolver(dict(cfg, backend=oracle.OracleBackend))
    U0 = _random_state(S, 7)
    S.setState(U0)
    S.boundary()
    got = S.getState().reshape(S.gridSize[2], S.gridSize[1], S.gridSize[0], -1)
    want = numpy_boundary(U0.reshape(got.shape), cfg["dim"], bc, _fixed_states(S), _flips(cfg["eqn"]))
    assert np.array_equal(got, want)


def test_extrapolation_is_exact_for_polynomials(hydrob200, oracle):
    cfg = dict(eqn="euler", dim=2, gridSize=[12, 10], backend=oracle.OracleBackend,
               boundary=dict(xmin="linear", xmax="linear", ymin="quadratic", ymax="quadratic"))
    S = hydrob200.FiniteVolumeSolver(cfg)
    j, i = np.meshgrid(np.arange(S.gridSize[1], dtype=float), np.arange(S.gridSize[0], dtype=float), indexing="ij")
    f = 3. + 2. * i + (1. + .5 * i) * (j * j - 4. * j)        # linear in i, quadratic in j: integers and halves, exact in double
    U = np.repeat(f.reshape(-1, 1), S.eqn.numStates, axis=1)
    garbage = U.copy().reshape(S.gridSize[1], S.gridSize[0], -1)
    garbage[:2] = garbage[-2:] = -77.
    garbage[:, :2] = garbage[:, -2:] = -77.
    S.setState(garbage.reshape(U.shape))
    S.boundary()
    assert np.array_equal(S.getState().reshape(U.shape), U)


def test_fixed_without_state_is_an_error(hydrob200, oracle):
    with pytest.raises(ValueError):
        hydrob200.FiniteVolumeSolver(dict(eqn="euler", dim=1, gridSize=[16], boundary=dict(xmin="fixed"), backend=oracle.OracleBackend))


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["double", "float"])
@pytest.mark.parametrize("name", list(CONFIGS))
def test_gpu_boundary_equals_oracle(hydrob200, oracle, name, precision):
    cfg, bc = CONFIGS[name]
    G = hydrob200.FiniteVolumeSolver(dict(cfg, precision=precision, strict_fp=True))
    R = hydrob200.FiniteVolumeSolver(dict(cfg, precision=precision, backend=oracle.OracleBackend))
    U0 = _random_state(R, 11).astype(np.float32 if precision == "float" else np.float64).astype(np.float64)
    for S in (G, R):
        S.setState(U0)
        S.boundary()
    assert np.array_equal(G.getState(), R.getState())
    assert G.backend.launch_count() >= cfg["dim"]


@pytest.mark.gpu
def test_gpu_set_fixed_boundary_rejects_other_faces(hydrob200):
    G = hydrob200.FiniteVolumeSolver(dict(eqn="euler", dim=2, gridSize=[16, 16], boundary=dict(ymax=LID)))
    with pytest.raises(Exception):
        G.backend.set_fixed_boundary(0, [1., 0., 0., 0., 1., 0.])
