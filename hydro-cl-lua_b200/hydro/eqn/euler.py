"""Compressible Euler equations: host side of hydro/eqn/euler.lua.

Reference: euler.lua:13-14 (numWaves 5, numIntStates 5), :166-171 (cons_t = rho, m[3], ETotal, ePot),
:190-197 (heatCapacityRatio 7/5, rhoMin 1e-7, PMin 1e-7); consFromPrim euler.cl:163-173.
Device functions: csrc/eqn_euler.cuh.
"""
from .eqn import Equation


class Euler(Equation):
    name = "euler"
    eqnId = 0
    numStates = 6
    numIntStates = 5
    numWaves = 5
    consVars = ("rho", "mx", "my", "mz", "ETotal", "ePot")
    guiVars = {"heatCapacityRatio": 7. / 5., "rhoMin": 1e-7, "PMin": 1e-7}

    @property
    def heatCapacityRatio(self):
        return self.vars["heatCapacityRatio"]

    def eqnParams(self):
        v = self.vars
        return [v["heatCapacityRatio"], v["rhoMin"], v["PMin"]]

    def consFromPrim(self, W):
        g = self.vars["heatCapacityRatio"]
        rho = W["rho"]
        vSq = W["vx"] * W["vx"] + W["vy"] * W["vy"] + W["vz"] * W["vz"]
        return dict(rho=rho, mx=W["vx"] * rho, my=W["vy"] * rho, mz=W["vz"] * rho,
                    ETotal=(rho * (.5 * vSq)) + (W["P"] / (g - 1.)), ePot=W["ePot"])
