"""Initial conditions of the numerical-relativity config (hydro/init/einstein.lua), as numpy formulas evaluated on the host.

applyInitCondCell (hydro/eqn/adm3d.cl:161-194) starts from alpha = 1, gamma_ll = identity (Cartesian coord_g_ll), K_ll = 0 and
runs the initial condition's code; a_l, d_lll and V_l are then derived on the device by the initDerivs kernel.
"""
import numpy as np

from .euler import InitCond


def _flat(x):
    one, z = np.ones_like(x), np.zeros_like(x)
    W = dict(alpha=one.copy())
    for s in ("xx", "xy", "xz", "yy", "yz", "zz"):
        W["gamma_ll_" + s] = one.copy() if s in ("xx", "yy", "zz") else z.copy()
        W["K_ll_" + s] = z.copy()
    return W


class GaugeWave(InitCond):
    """'testbed - gauge wave' (init/einstein.lua:1485-1507): H = 1 + A sin(2 pi x / d); alpha = sqrt(H), gamma_xx = H,
    K_xx = -pi A / d cos(theta) / alpha.  (The reference leaves its +-.5 domain / periodic boundary lines commented out; the
    config sets them explicitly: SURVEY 8d C5.)"""
    name = "testbed - gauge wave"
    guiVars = {"A": .1, "d": 1.}

    def prims(self, x, y, z, solver):
        v = self.vars
        W = _flat(x)
        theta = 2. * np.pi / v["d"] * (x - 0.)
        H = 1. + v["A"] * np.sin(theta)
        W["alpha"] = np.sqrt(H)
        W["gamma_ll_xx"] = H
        W["K_ll_xx"] = -np.pi * v["A"] / v["d"] * np.cos(theta) / W["alpha"]
        return W


class AlcubierreWarpBubble(InitCond):
    """'Alcubierre warp bubble' (init/einstein.lua:493-545), R = .5, sigma = 8, speed = .1; with useShift = 'none' the shift it
    defines is not part of the state."""
    name = "Alcubierre warp bubble"
    guiVars = {"R": .5, "sigma": 8., "speed": .1}

    def prims(self, x, y, z, solver):
        v = self.vars
        W = _flat(x)
        R, sigma, v_s = v["R"], v["sigma"], v["speed"]
        r_s = np.sqrt(x * x + y * y + z * z)

        def dtanh(a):
            c = .5 * (np.exp(a) + np.exp(-a))
            return 1. / (c * c)
        fdenom = 2 * np.tanh(sigma * R)
        with np.errstate(divide="ignore", invalid="ignore"):
            scale = sigma * (dtanh(sigma * (r_s + R)) - dtanh(sigma * (r_s - R))) / fdenom
            dx_f = [c * (1. / r_s) * scale for c in (x, y, z)]
        W["K_ll_xx"] = -v_s * dx_f[0] / 1.
        W["K_ll_xy"] = -v_s * dx_f[1] / 2.
        W["K_ll_xz"] = -v_s * dx_f[2] / 2.
        return W


initConds = {c.name: c for c in (GaugeWave, AlcubierreWarpBubble)}
