"""Explicit integrators of hydro/int/all.lua as Butcher-style (alphas, betas) tableaux.

Reference: hydro/int/all.lua:11-173 (tables), hydro/int/rk.lua:47-167 (how they are consumed:
stage i: U = sum_k alpha[i][k] U_k + dt * sum_k beta[i][k] L(U_k), alpha terms first, k ascending),
hydro/int/fe.lua:33-49 (forward Euler).  Implicit integrators (icn, be*) are out of scope.
"""

integrators = {
    "forward Euler": None,
    "Runge-Kutta 2": ([[1, 0], [1, 0]], [[.5, 0], [0, 1]]),
    "Runge-Kutta 2 Heun": ([[1, 0], [1, 0]], [[1, 0], [.5, .5]]),
    "Runge-Kutta 2 Ralston": ([[1, 0], [1, 0]], [[2. / 3., 0], [1. / 4., 3. / 4.]]),
    "Runge-Kutta 3": ([[1, 0, 0], [1, 0, 0], [1, 0, 0]],
                      [[.5, 0, 0], [-1, 2, 0], [1. / 6., 2. / 6., 1. / 6.]]),
    "Runge-Kutta 4": ([[1, 0, 0, 0]] * 4,
                      [[.5, 0, 0, 0], [0, .5, 0, 0], [0, 0, 1, 0], [1. / 6., 2. / 6., 2. / 6., 1. / 6.]]),
    "Runge-Kutta 4, 3/8ths rule": ([[1, 0, 0, 0]] * 4,
                                   [[1. / 3., 0, 0, 0], [-1. / 3., 0, 0, 0], [1, -1, 1, 0],
                                    [1. / 8., 3. / 8., 3. / 8., 1. / 8.]]),
    "Runge-Kutta 2, TVD": ([[1, 0], [.5, .5]], [[1, 0], [0, .5]]),
    "Runge-Kutta 2, non-TVD": ([[1, 0], [1, 0]], [[-20, 0], [41. / 40., -1. / 40.]]),
    "Runge-Kutta 3, TVD": ([[1, 0, 0], [3 / 4, 1 / 4, 0], [1 / 3, 0, 2 / 3]],
                           [[1, 0, 0], [0, 1 / 4, 0], [0, 0, 2 / 3]]),
    "Runge-Kutta 4, TVD": (
        [[1, 0, 0, 0],
         [649. / 1600., 951. / 1600., 0, 0],
         [53989. / 2500000., 4806213. / 20000000., 23619. / 32000., 0],
         [1. / 5., 6127. / 30000., 7873. / 30000., 1. / 3.]],
        [[.5, 0, 0, 0],
         [-10890423. / 25193600., 5000. / 7873, 0, 0],
         [-102261. / 5000000., -5121. / 20000., 7873. / 10000., 0],
         [1. / 10., 1. / 6., 0, 1. / 6.]]),
    "Runge-Kutta 4, non-TVD": ([[1, 0, 0, 0], [1, 0, 0, 0], [1, 0, 0, 0], [-1. / 3., 1. / 3., 2. / 3., 1. / 3.]],
                               [[.5, 0, 0, 0], [0, .5, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1. / 6.]]),
}

integratorNames = list(integrators.keys())


def tableau(name):
    """-> (order, alphas_flat16, betas_flat16); order 0 = forward Euler."""
    if name not in integrators:
        raise KeyError("unknown integrator %r (have: %s)" % (name, ", ".join(integratorNames)))
    tab = integrators[name]
    a = [0.0] * 16
    b = [0.0] * 16
    if tab is None:
        return 0, a, b
    alphas, betas = tab
    order = len(alphas)
    for i in range(order):
        for k in range(order):
            a[i * order + k] = float(alphas[i][k])
            b[i * order + k] = float(betas[i][k])
    return order, a, b
