"""The codegen seam (VERDICT r01 'g1'): an equation's device functions handed over as SOURCE at run time and compiled by NVRTC into the
fused marching kernels -- where the reference's eqn:initCodeModules output lands (hydro/eqn/eqn.lua:506-513,577-635;
hydro/solver/solverbase.lua:1686-1699).

  * hb_eqn_euler.cuh's own text fed back under its own name reproduces the ahead-of-time build bit for bit (production and strict);
  * a RENAMED copy with another eqnId -- "a fourth equation", which takes the generic Roe path through the plug-in contract instead of the
    hand-tuned Euler forms -- runs without rebuilding libhydrob200.so: its strict build is bit-identical to the strict Euler build (the
    same literal functions), its production build stays within 1e-12 of the oracle;
  * the same in 3-D, and for ideal MHD's 8-variable system;
  * a header with an error returns NVRTC's log, not a crash."""
import os

import numpy as np
import pytest

from cases import CASES

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "hydro-cl-lua_b200", "csrc")


def run(hydrob200, cfg, n, **kw):
    S = hydrob200.FiniteVolumeSolver(dict(cfg, **kw))
    for _ in range(n):
        S.update()
    return S.getState(), S.t, S


def euler_header():
    return open(os.path.join(CSRC, "hb_eqn_euler.cuh")).read()


def renamed(text, old, new, eqnId):
    out = text.replace("struct %s {" % old, "struct %s {" % new).replace("static constexpr int eqnId = ", "static constexpr int eqnId = %d + 0 * " % eqnId)
    assert out != text
    return out


@pytest.mark.parametrize("case", ["C2_kh_rk4tvd_minmod", "C4_sphere_rk4"])
@pytest.mark.parametrize("strict", [False, True])
def test_same_source_gives_the_same_kernels(hydrob200, case, strict):
    cfg, n = CASES[case]
    a, ta, A = run(hydrob200, cfg, n, strict_fp=strict)
    b, tb, B = run(hydrob200, cfg, n, strict_fp=strict, eqnSource=dict(name="hb_eqn_euler.cuh", src=euler_header(), type="hb::Euler"))
    assert "eqn=100" in B.backend.describe() and "fv_march" in B.backend.describe()
    assert ta == tb
    assert np.array_equal(a, b)


@pytest.mark.parametrize("case", ["C2_kh_rk4tvd_superbee", "C4_sphere_rk4_mirror_periodic"])
def test_a_fourth_equation_without_rebuilding(hydrob200, oracle, case):
    cfg, n = CASES[case]
    src = renamed(euler_header(), "Euler", "UserGas", 7)
    user = dict(name="user_gas.cuh", src="#pragma once\n#include \"hb_math.cuh\"\n" + src, type="hb::UserGas")
    ref, tref, _ = run(hydrob200, cfg, n, backend=oracle.OracleBackend)
    got, tgot, S = run(hydrob200, cfg, n, strict_fp=True, eqnSource=user)
    assert "eqn=107" in S.backend.describe()
    assert tgot == tref
    assert np.array_equal(got, ref)                       # strict: the literal functions, through the generic Roe path
    got, tgot, _ = run(hydrob200, cfg, n, eqnSource=user)
    scale = np.abs(ref[..., :5]).max(axis=(0, 1, 2))
    assert (np.abs(got[..., :5] - ref[..., :5]).max(axis=(0, 1, 2)) <= 1e-12 * np.maximum(scale, scale.max() * 1e-3)).all()


def test_mhd_from_source(hydrob200):
    cfg, n = CASES["C3_ot_rk3tvd"]
    src = open(os.path.join(CSRC, "hb_eqn_mhd.cuh")).read()
    a, ta, _ = run(hydrob200, cfg, n, strict_fp=True)
    b, tb, _ = run(hydrob200, cfg, n, strict_fp=True, eqnSource=dict(name="hb_eqn_mhd.cuh", src=src, type="hb::MHD"))
    assert ta == tb and np.array_equal(a, b)


def test_compile_error_is_reported(hydrob200):
    from importlib import import_module
    hb = import_module("hydro-cl-lua_b200._lib")
    cfg, _ = CASES["C2_kh_rk4tvd_minmod"]
    bad = euler_header().replace("W.P = calc_P(s, U);", "W.P = calc_P(s, U) + undeclared_symbol;")
    with pytest.raises(hb.HydroB200Error) as e:
        hydrob200.FiniteVolumeSolver(dict(cfg, eqnSource=dict(name="hb_eqn_euler.cuh", src=bad, type="hb::Euler")))
    assert "undeclared_symbol" in str(e.value)
    # and a configuration outside the marching kernels is refused, not silently run on another path
    with pytest.raises(hb.HydroB200Error):
        hydrob200.FiniteVolumeSolver(dict(eqn="euler", dim=1, gridSize=[64], initCond="Sod", usePLM="plm cons", slopeLimiter="minmod",
                                          eqnSource=dict(name="hb_eqn_euler.cuh", src=euler_header(), type="hb::Euler")))
