"""Initial conditions used by the five BASELINE configs (hydro/init/euler.lua), as numpy formulas.

These are *synthetic inputs*: they are evaluated once on the host in double precision and handed,
as AoS cons_t records, to whichever backend runs (CUDA product or CPU oracle), so transcendental
differences between libms can never show up as a parity difference.  The reference evaluates them
in its ``applyInitCond`` OpenCL kernel (hydro/eqn/eqn.lua:622-632, euler.cl:236-269, mhd.cl:251-287).

Each initial condition is a function ``f(x, y, z, solver) -> dict(rho, vx, vy, vz, P, Bx, By, Bz, ePot)``
plus class-level overrides (solverVars, boundary, mins/maxs) exactly where the reference sets them.
"""
import math

import numpy as np


class InitCond:
    name = None
    solverVars = {}
    mins = None
    maxs = None
    boundary = None     # name applied by solver:setBoundaryMethods(...) in getInitCondCode
    guiVars = {}

    def __init__(self, args=None):
        self.args = dict(args or {})
        self.vars = dict(self.guiVars)
        for k, v in self.args.items():
            self.vars[k] = v

    def prims(self, x, y, z, solver):
        raise NotImplementedError

    def exactSolution(self, t, x, solver):
        raise NotImplementedError("no exact solution for " + str(self.name))


def _zeros_like(x):
    return np.zeros_like(x)


class Sod(InitCond):
    """RiemannProblem 'Sod' (init/euler.lua:57-120, 743-752): gamma overridden to 5/3."""
    name = "Sod"
    solverVars = {"heatCapacityRatio": 5. / 3.}
    guiVars = {"rhoL": 1., "PL": 1., "rhoR": .125, "PR": .1}

    def prims(self, x, y, z, solver):
        mids = [.5 * (solver.initCondMins[i] + solver.initCondMaxs[i]) for i in range(3)]
        dim = self.args.get("dim", solver.dim)
        lhs = np.ones_like(x, dtype=bool)
        for i, c in enumerate((x, y, z)[:dim]):
            lhs &= c < mids[i]
        v = self.vars
        z0 = _zeros_like(x)
        return dict(rho=np.where(lhs, v["rhoL"], v["rhoR"]), vx=z0, vy=z0, vz=z0,
                    P=np.where(lhs, v["PL"], v["PR"]), Bx=z0, By=z0, Bz=z0, ePot=z0)

    def exactSolution(self, t, x, solver):
        """Exact Riemann solution, init/euler.lua:124-258 (Newton iteration written as in the reference)."""
        v = self.vars
        rhoL, rhoR, PL, PR = v["rhoL"], v["rhoR"], v["PL"], v["PR"]
        vL = vR = 0.
        gamma = solver.heatCapacityRatio
        muSq = (gamma - 1) / (gamma + 1)
        CsL = math.sqrt(gamma * PL / rhoL)
        CsR = math.sqrt(gamma * PR / rhoR)
        s75 = math.sqrt(0.75)
        P3 = .5 * (PL + PR)
        for _ in range(1000):
            f = ((((-2 * CsL) * (1 - ((P3 / PL) ** ((-1 + gamma) / (2 * gamma))))) / (CsR * (-1 + gamma)))
                 + ((-1 + (P3 / PR)) * ((0.75 / (gamma * (0.25 + (P3 / PR)))) ** 0.5)))
            A = PL ** ((1 - gamma) / (2 * gamma))
            q = math.sqrt(P3 + 0.25 * PR)
            df = ((-((((((1.5 * s75 * CsR * PR * (gamma ** 1.5)) - ((0.75 ** 1.5) * CsR * PR * math.sqrt(gamma)))
                        - ((0.75 ** 1.5) * CsR * (gamma ** 2.5) * PR)) - (0.5 * P3 * s75 * CsR * math.sqrt(gamma)))
                      - (0.5 * P3 * s75 * CsR * (gamma ** 2.5)))
                     + (((P3 * s75 * CsR * (gamma ** 1.5))
                         - (0.25 * A * CsL * (P3 ** (((-1) - gamma) / (2 * gamma))) * (PR ** 1.5) * q))
                        - (A * CsL * (P3 ** ((-(1 - gamma)) / (2 * gamma))) * math.sqrt(PR) * q))
                     + (0.5 * A * CsL * (P3 ** (((-1) - gamma) / (2 * gamma))) * (PR ** 1.5) * gamma * q)
                     + (((2 * A * CsL * (P3 ** ((-(1 - gamma)) / (2 * gamma))) * math.sqrt(PR) * gamma * q)
                         - (0.25 * A * CsL * (P3 ** (((-1) - gamma) / (2 * gamma))) * (gamma ** 2) * (PR ** 1.5) * q))
                        - (A * CsL * (P3 ** ((-(1 - gamma)) / (2 * gamma))) * (gamma ** 2) * math.sqrt(PR) * q))))
                  / (math.sqrt(PR) * CsR * ((P3 + (0.25 * PR)) ** 1.5) * gamma * ((1 - (2 * gamma)) + (gamma ** 2))))
            dP3 = -f / df
            if abs(dP3) <= 1e-16:
                break
            if not math.isfinite(dP3):
                raise FloatingPointError("delta is not finite")
            P3 = P3 + dP3
        P4 = P3
        rho3 = rhoL * (P3 / PL) ** (1 / gamma)
        v3 = vR + 2 * CsL / (gamma - 1) * (1 - (P3 / PL) ** ((gamma - 1) / (2 * gamma)))
        v4 = v3
        rho4 = rhoR * (P4 + muSq * PR) / (PR + muSq * P4)
        vshock = v4 * rho4 / (rho4 - rhoR)
        vtail = CsL - v4 / (1 - muSq)
        s1, s2, s3, s4 = -CsL, -vtail, v3, vshock
        x = np.asarray(x, dtype=np.float64)
        with np.errstate(divide="ignore", invalid="ignore"):
            xi = x / t
            fan = -muSq * (x / (CsL * t)) + (1 - muSq)
            rho = np.where(xi < s1, rhoL, np.where(xi < s2, rhoL * fan ** (2 / (gamma - 1)),
                  np.where(xi < s3, rho3, np.where(xi < s4, rho4, rhoR))))
            vx = np.where(xi < s1, vL, np.where(xi < s2, (1 - muSq) * (x / t + CsL),
                 np.where(xi < s3, v3, np.where(xi < s4, v4, vR))))
            P = np.where(xi < s1, PL, np.where(xi < s2, PL * fan ** (2 * gamma / (gamma - 1)),
                np.where(xi < s3, P3, np.where(xi < s4, P4, PR))))
        EInt = P / (gamma - 1)
        EKin = .5 * rho * (vx * vx)
        return rho, rho * vx, 0 * rho, 0 * rho, EKin + EInt


class AdvectWave(InitCond):
    """'advect wave' (init/euler.lua:641-680): domain [0,1], periodic, gamma 7/5."""
    name = "advect wave"
    mins = (0., 0., 0.)
    maxs = (1., 1., 1.)
    boundary = "periodic"
    solverVars = {"heatCapacityRatio": 7. / 5.}
    guiVars = {"rho0": 1., "rho1": 3.2e-1, "v0x": 1., "P0": 1.}

    def _rho(self, t, x, solver):
        v = self.vars
        k0 = 2 * math.pi / (solver.maxs[0] - solver.mins[0])
        return v["rho0"] + v["rho1"] * np.sin(k0 * (x - solver.mins[0] - v["v0x"] * t))

    def prims(self, x, y, z, solver):
        z0 = _zeros_like(x)
        return dict(rho=self._rho(0., x, solver), vx=z0 + self.vars["v0x"], vy=z0, vz=z0,
                    P=z0 + self.vars["P0"], Bx=z0, By=z0, Bz=z0, ePot=z0)

    def exactSolution(self, t, x, solver):
        v = self.vars
        rho = self._rho(t, np.asarray(x, dtype=np.float64), solver)
        vx = v["v0x"]
        ETotal = v["P0"] / (solver.heatCapacityRatio - 1) + 0.5 * rho * (vx ** 2)
        return rho, rho * vx, 0 * rho, 0 * rho, ETotal


class KelvinHelmholtz(InitCond):
    """'Kelvin-Helmholtz' (init/euler.lua:1640-1744), periodic.  noiseAmplitude defaults to 0 here
    (the reference's 1e-2 consumes host math.random() values, init/init.lua:211-219: not reproducible)."""
    name = "Kelvin-Helmholtz"
    boundary = "periodic"
    guiVars = {"rhoInside": 2., "rhoOutside": 1., "amplitude": 1e-2, "noiseAmplitude": 0.,
               "backgroundPressure": 2.5, "frequency": 2., "thickness": .025, "velInside": -.5, "velOutside": .5}

    def prims(self, x, y, z, solver):
        v = self.vars
        if v["noiseAmplitude"] != 0:
            raise ValueError("noiseAmplitude != 0 needs the reference's host RNG stream; use 0")
        sliceAxis = self.args.get("sliceAxis", 2) - 1    # 1-based in the reference
        moveAxis = self.args.get("moveAxis", 1) - 1
        coords = (x, y, z)
        mins, maxs = solver.mins, solver.maxs
        yq1 = mins[sliceAxis] * .75 + maxs[sliceAxis] * .25
        yq2 = mins[sliceAxis] * .25 + maxs[sliceAxis] * .75
        xs = coords[sliceAxis]
        inside = (.5 + .5 * np.tanh((xs - yq1) / v["thickness"])) - (.5 + .5 * np.tanh((xs - yq2) / v["thickness"]))
        theta = v["frequency"] * 2. * math.pi
        theta = theta + 0 * x
        for i in range(solver.dim):
            if i != sliceAxis:
                theta = theta * ((coords[i] - mins[i]) / (maxs[i] - mins[i]))
        noise = (maxs[0] - mins[0]) * v["amplitude"]
        rho = inside * v["rhoInside"] + (1. - inside) * v["rhoOutside"]
        vel = [np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)]
        if solver.dim >= 2:
            perpAxis = 1 if solver.dim == 2 else 2
            vel[perpAxis] = np.sin(theta) * noise
        vel[moveAxis] = vel[moveAxis] + (inside * v["velInside"] + (1. - inside) * v["velOutside"])
        z0 = _zeros_like(x)
        return dict(rho=rho, vx=vel[0], vy=vel[1], vz=vel[2], P=z0 + v["backgroundPressure"],
                    Bx=z0, By=z0, Bz=z0, ePot=z0)


class OrszagTang(InitCond):
    """'Orszag-Tang' (init/euler.lua:961-980), CSUN variant, gamma 5/3, periodic; keeps the reference's
    B.y = B0 sin(2 pi (x + .5)) (no x/2) at :977."""
    name = "Orszag-Tang"
    boundary = "periodic"
    solverVars = {"heatCapacityRatio": 5. / 3.}

    def prims(self, x, y, z, solver):
        g = solver.heatCapacityRatio
        B0 = 1. / math.sqrt(4. * math.pi)
        z0 = _zeros_like(x)
        return dict(rho=z0 + g * g,
                    vx=-np.sin(2. * math.pi * (y * .5 + .5)), vy=np.sin(2. * math.pi * (x * .5 + .5)), vz=z0,
                    P=z0 + g,
                    Bx=-B0 * np.sin(2. * math.pi * (y * .5 + .5)), By=B0 * np.sin(2. * math.pi * (x + .5)), Bz=z0,
                    ePot=z0)


class Sphere(InitCond):
    """'sphere' blast (init/euler.lua:1274-1296). Note the reference's PInside default reads args.rhoInside."""
    name = "sphere"
    guiVars = {"radius": .5, "rhoInside": 1., "PInside": 1., "rhoOutside": .01, "POutside": .01}

    def prims(self, x, y, z, solver):
        v = self.vars
        rSq = x * x + y * y + z * z
        inside = rSq < v["radius"] * v["radius"]
        z0 = _zeros_like(x)
        return dict(rho=np.where(inside, v["rhoInside"], v["rhoOutside"]), vx=z0, vy=z0, vz=z0,
                    P=np.where(inside, v["PInside"], v["POutside"]), Bx=z0, By=z0, Bz=z0, ePot=z0)


class BrioWu(InitCond):
    """'Brio-Wu' MHD shock tube (init/euler.lua, RiemannProblem): gamma 2, Bx=.75, By=+-1."""
    name = "Brio-Wu"
    solverVars = {"heatCapacityRatio": 2.}

    def prims(self, x, y, z, solver):
        mid = .5 * (solver.initCondMins[0] + solver.initCondMaxs[0])
        lhs = x < mid
        z0 = _zeros_like(x)
        return dict(rho=np.where(lhs, 1., .125), vx=z0, vy=z0, vz=z0, P=np.where(lhs, 1., .1),
                    Bx=z0 + .75, By=np.where(lhs, 1., -1.), Bz=z0, ePot=z0)


initConds = {c.name: c for c in (Sod, AdvectWave, KelvinHelmholtz, OrszagTang, Sphere, BrioWu)}
initCondNames = list(initConds.keys())
