"""State dumps in the reference's FITS layout (hydro/solver/gridsolver.lua:1410-1470): structure of the file against the
FITS standard, the reference's element order, and the save -> load round trip through the solver."""
import os
import struct

import numpy as np


def test_layout_and_standard_conformance(hydrob200, tmp_path):
    fits = __import__("importlib").import_module("hydro-cl-lua_b200.hydro.fits")
    w, h, d, c = 5, 4, 3, 2
    U = np.arange(d * h * w * c, dtype=np.float64).reshape(d, h, w, c) + .25
    fn = str(tmp_path / "a.fits")
    fits.write_image(fn, fits.state_to_image(U))
    raw = open(fn, "rb").read()
    assert len(raw) % 2880 == 0
    cards = [raw[n:n + 80].decode("ascii") for n in range(0, 2880, 80)]
    assert cards[0].startswith("SIMPLE  =                    T")
    assert cards[1].startswith("BITPIX  =                  -64")
    assert [cards[n][:8] for n in range(2, 6)] == ["NAXIS   ", "NAXIS1  ", "NAXIS2  ", "NAXIS3  "]
    assert int(cards[3][10:30]) == w * d and int(cards[4][10:30]) == h and int(cards[5][10:30]) == c
    assert any(x.startswith("END") for x in cards)
    # gridsolver.lua:1429: element (ch, i, j, k) at i + width * (k + depth * (j + height * ch)), big-endian doubles after the header
    for ch, i, j, k in [(0, 0, 0, 0), (1, 4, 3, 2), (0, 2, 1, 1), (1, 0, 2, 1)]:
        off = 2880 + 8 * (i + w * (k + d * (j + h * ch)))
        assert struct.unpack(">d", raw[off:off + 8])[0] == U[k, j, i, ch]
    assert np.array_equal(fits.image_to_state(fits.read_image(fn), (w, h, d)), U)


def test_float_dump(hydrob200, tmp_path):
    fits = __import__("importlib").import_module("hydro-cl-lua_b200.hydro.fits")
    U = np.linspace(0, 1, 2 * 3 * 4 * 6).reshape(2, 3, 4, 6)
    fn = str(tmp_path / "f.fits")
    fits.write_image(fn, fits.state_to_image(U, np.float32))
    assert open(fn, "rb").read(2880).decode("ascii")[80:110].strip().endswith("-32")
    assert np.array_equal(fits.image_to_state(fits.read_image(fn), (4, 3, 2)), U.astype(np.float32).astype(np.float64))


def test_solver_save_load_round_trip(hydrob200, oracle, tmp_path):
    cfg = dict(eqn="euler", dim=3, gridSize=[10, 8, 6], mins=[-2] * 3, maxs=[2] * 3, initCond="sphere", usePLM="plm cons",
               slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.1, backend=oracle.OracleBackend)
    A = hydrob200.FiniteVolumeSolver(cfg)
    A.update()
    fn = A.save(str(tmp_path / "dump"))
    assert os.path.basename(fn) == "dump_UBuf.fits"
    B = hydrob200.FiniteVolumeSolver(cfg)
    B.load(str(tmp_path / "dump"))
    assert np.array_equal(A.getState(), B.getState())
    B.t = A.t
    B.backend.set_t(A.t)
    A.update()
    B.update()
    assert np.array_equal(A.getState(), B.getState())


def test_compare_dump_tool(hydrob200, oracle, tmp_path):
    """tools/compare_dump.py: a dump of the same configuration compares clean, a perturbed one is reported (SURVEY 8c hook)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("compare_dump", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "compare_dump.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    from cases import CASES
    cfg, _ = CASES["C2_kh_rk4tvd_minmod"]
    S = hydrob200.FiniteVolumeSolver(dict(cfg, backend=oracle.OracleBackend))
    for _ in range(3):
        S.update()
    fn = S.save(str(tmp_path / "ref"))
    R = hydrob200.FiniteVolumeSolver(dict(cfg, backend=oracle.OracleBackend))
    for _ in range(3):
        R.update()
    assert tool.compare(R, fn) == 0
    U = S.getState()
    U[0, 10, 10, 0] *= 1. + 1e-9
    S.saveBuffer(U, str(tmp_path / "bad_UBuf"))
    assert tool.compare(R, str(tmp_path / "bad_UBuf.fits")) == 1
