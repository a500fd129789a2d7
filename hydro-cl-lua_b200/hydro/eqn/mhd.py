"""Ideal MHD (Stone et al. 2008 / Athena eigensystem): host side of hydro/eqn/mhd.lua.

Reference: mhd.lua:16-17 (numWaves 7, numIntStates 8), :76-83 (cons_t = rho, m[3], ETotal, B[3], psi, ePot),
:209-217 (heatCapacityRatio 2, mu0 = vacuum permeability), :234 (coulomb = sqrt(kg*m/mu0) so that
solver->mu0 / unit_kg_m_per_C2 ~= 1, math.cl:270); consFromPrim mhd.cl:184-201.
Parity contract: the NoDiv / selfgrav ops the reference installs (mhd.lua:113-123) are disabled.
Device functions: csrc/eqn_mhd.cuh.
"""
import math

from .eqn import Equation

vacuumPermeability_in_kg_m_per_C2 = 1.2566370621219e-6    # hydro/constants.lua:10


class MHD(Equation):
    name = "mhd"
    eqnId = 1
    numStates = 10
    numIntStates = 8
    numWaves = 7
    consVars = ("rho", "mx", "my", "mz", "ETotal", "Bx", "By", "Bz", "psi", "ePot")
    guiVars = {"heatCapacityRatio": 2., "mu0": vacuumPermeability_in_kg_m_per_C2}

    def __init__(self, solver, args=None):
        super().__init__(solver, args)
        v = self.vars
        v["coulomb"] = math.sqrt(v["kilogram"] * v["meter"] / v["mu0"])   # mhd.lua:234

    @property
    def heatCapacityRatio(self):
        return self.vars["heatCapacityRatio"]

    @property
    def mu0_eff(self):
        """solver->mu0 / unit_kg_m_per_C2 with unit_kg_m_per_C2 = kg * m / (C * C) (math.cl:252,270)."""
        v = self.vars
        return v["mu0"] / (v["kilogram"] * v["meter"] / (v["coulomb"] * v["coulomb"]))

    def eqnParams(self):
        """{gamma, mu0 / unit_kg_m_per_C2, sign of the left-eigenvector entry l23}.  The reference's eigen_leftTransform has
        l23 = +.5 betaZ (mhd.cl:621); Stone et al. 2008 -- and R L = I -- need -.5 betaZ.  The reference's sign is the parity contract and
        the default; eqnArgs = {stone2008_l23 = true} selects the corrected one (tests/test_mhd_alfven.py shows what it changes)."""
        return [self.vars["heatCapacityRatio"], self.mu0_eff, -1. if self.args.get("stone2008_l23") else 1.]

    def consFromPrim(self, W):
        g = self.vars["heatCapacityRatio"]
        rho = W["rho"]
        vSq = W["vx"] * W["vx"] + W["vy"] * W["vy"] + W["vz"] * W["vz"]
        BSq = W["Bx"] * W["Bx"] + W["By"] * W["By"] + W["Bz"] * W["Bz"]
        EKin = .5 * rho * vSq
        EMag = .5 * BSq / self.mu0_eff
        EInt = W["P"] / (g - 1.)
        return dict(rho=rho, mx=W["vx"] * rho, my=W["vy"] * rho, mz=W["vz"] * rho, ETotal=EInt + EKin + EMag,
                    Bx=W["Bx"], By=W["By"], Bz=W["Bz"], psi=0 * rho, ePot=W["ePot"])
