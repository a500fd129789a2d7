"""The static plan hb_fv_update runs for a Butcher tableau (csrc/hb_fv.cu buildPlan / foldFinalStage, exported host-only as hb_rk_plan):
executed on scalars with numpy doubles and compared, BITWISE, with the direct evaluation of hydro/int/rk.lua:91-165

    U^(i+1) = ((0 + sum_k alpha[i][k] U^k) + sum_k (beta[i][k] dt) L(U^k)),   alpha terms first, k ascending

for every tableau of hydro/int/all.lua, with and without the folded last stage.  No GPU: the plan is host logic."""
import ctypes as C
import re
from importlib import import_module

import numpy as np
import pytest

int_all = import_module("hydro-cl-lua_b200.hydro.int.all")
NAMES = [n for n in int_all.integratorNames if int_all.integrators[n] is not None]


def plan_of(name, fold):
    import hydrob200  # noqa: F401  (puts the package on the path)
    lib = import_module("hydro-cl-lua_b200._lib")
    L = lib.lib()
    order, alphas, betas = int_all.tableau(name)
    A = (C.c_double * 16)(*alphas)
    B = (C.c_double * 16)(*betas)
    out = C.create_string_buffer(8192)
    lib.check(L.hb_rk_plan(order, A, B, 1 if fold else 0, out, 8192))
    lines = out.value.decode().strip().split("\n")
    head = dict(kv.split("=") for kv in lines[0].split())
    stages = []
    for ln in lines[1:]:
        kv = dict(t.split("=", 1) for t in ln.split()[2:])
        terms = lambda s: [(int(a.split(":")[0]), float(a.split(":")[1])) for a in s.split(",")] if s else []
        stages.append(dict(uIn=int(kv["in"]), uOut=int(kv["out"]), lOut=int(kv["lout"]), computeL=int(kv["computeL"]), betaSelf=float(kv["betaSelf"]),
                           operands=int(kv["operands"]), alpha=terms(kv["alpha"]), beta=terms(kv["beta"]), accOut=int(kv["accOut"]), accIn=int(kv["accIn"]),
                           accCoef=float(kv["accCoef"]), accBetaSelf=float(kv["accBetaSelf"])))
    return int(head["nU"]), int(head["nL"]), int(head["folded"]), stages, order, alphas, betas


def rhs(u):
    return np.sin(3. * u) - u * u * .25 + .1          # any nonlinear L(U)


def run_plan(nU, nL, stages, U0, dt):
    U = [None] * nU
    Lb = [None] * nL
    U[0] = U0.copy()
    for s in stages:
        own = U[s["uIn"]]
        Lval = rhs(own) if s["computeL"] else None
        if s["lOut"] >= 0:
            Lb[s["lOut"]] = Lval
        acc = None
        if s["accOut"] >= 0:           # the running sum is formed from the stage's operands BEFORE any buffer of this stage is written
            acc = (0. + own * s["accCoef"]) if s["accIn"] < 0 else U[s["accIn"]].copy()
            acc = acc + Lval * (s["accBetaSelf"] * dt)
        v = np.zeros_like(U0)
        for k, c in s["alpha"]:
            v = v + U[k] * c
        for k, c in s["beta"]:
            v = v + Lb[k] * (c * dt)
        if s["computeL"]:
            v = v + Lval * (s["betaSelf"] * dt)
        U[s["uOut"]] = v
        if acc is not None:
            U[s["accOut"]] = acc
    return U[stages[-1]["uOut"]]


def run_direct(order, alphas, betas, U0, dt):
    Us, Ls = [U0.copy()], []
    for i in range(order):
        Ls.append(rhs(Us[i]))
        v = np.zeros_like(U0)
        for k in range(i + 1):
            if alphas[i * order + k] != 0:
                v = v + Us[k] * alphas[i * order + k]
        for k in range(i + 1):
            if betas[i * order + k] != 0:
                v = v + Ls[k] * (betas[i * order + k] * dt)
        Us.append(v)
    return Us[-1]


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("fold", [False, True])
def test_plan_reproduces_the_tableau_bitwise(name, fold):
    nU, nL, folded, stages, order, alphas, betas = plan_of(name, fold)
    rng = np.random.default_rng(5)
    U0 = rng.uniform(-1., 1., 257)
    dt = .0371
    got = run_plan(nU, nL, stages, U0, dt)
    ref = run_direct(order, alphas, betas, U0, dt)
    assert np.array_equal(got, ref), (name, fold, np.abs(got - ref).max())
    assert all(s["operands"] <= 8 for s in stages)
    if not fold:
        assert not folded


def test_classic_rk4_folds_to_one_running_sum():
    nU, nL, folded, stages, *_ = plan_of("Runge-Kutta 4", True)
    assert folded == 1 and nL == 0 and nU == 4
    assert [s["operands"] for s in stages] == [0, 2, 2, 1]          # every stage fits the tallest marching tile (<= 2 staged operands)
    assert all(s["lOut"] < 0 for s in stages)
    nU0, nL0, folded0, stages0, *_ = plan_of("Runge-Kutta 4", False)
    assert folded0 == 0 and (nU0, nL0) == (3, 3) and [s["operands"] for s in stages0] == [0, 1, 1, 4]


@pytest.mark.parametrize("name", ["Runge-Kutta 4, TVD", "Runge-Kutta 3, TVD", "Runge-Kutta 4, 3/8ths rule", "Runge-Kutta 3", "Runge-Kutta 2"])
def test_other_tableaux_are_left_alone(name):
    """Dense alphas, a last stage that is not `U^0 + every earlier L`, or fewer than three stages: not folded."""
    assert plan_of(name, True)[2] == 0
