"""ADM Bona-Masso 3-D (first-order hyperbolic 3+1 numerical relativity): host side of hydro/eqn/adm3d.lua.

Reference: adm3d.lua:58-163 (consVars: alpha, gamma_ll, a_l, d_lll, K_ll, V_l integrated = 37 reals; rho, S_u, S_ll, H, M_u
carried = 14 reals), :124-143 (noZeroRowsInFlux -> numWaves = 13), :21 (roeUseFluxFromCons = false), :196-227 (guiVars),
hydro/eqn/einstein.lua:37-49 (f_eqn options).  useShift = 'none' only (the configs' setting).
Device functions: csrc/hb_eqn_adm3d.cuh.
"""
import numpy as np

from .eqn import Equation

SYM = ("xx", "xy", "xz", "yy", "yz", "zz")
F_EQN = ("2/alpha", "1 + 1/alpha^2", "1", "0", ".49", ".5", "1.5", "1.69")     # hydro/eqn/einstein.lua:42-48


class ADM3D(Equation):
    name = "adm3d"
    eqnId = 2
    numStates = 51
    numIntStates = 37
    numWaves = 13
    roeUseFluxFromCons = False
    consVars = (("alpha",) + tuple("gamma_ll_" + s for s in SYM) + tuple("a_l_" + x for x in "xyz")
                + tuple("d_lll_%s_%s" % (k, s) for k in "xyz" for s in SYM) + tuple("K_ll_" + s for s in SYM)
                + tuple("V_l_" + x for x in "xyz") + ("rho",) + tuple("S_u_" + x for x in "xyz") + tuple("S_ll_" + s for s in SYM)
                + ("H",) + tuple("M_u_" + x for x in "xyz"))
    guiVars = {"f_eqn": "2/alpha", "a_convCoeff": 0., "d_convCoeff": 0., "V_convCoeff": 10.,
               "K_ll_srcStressCoeff": 0., "K_ll_srcHCoeff": 0., "gamma_ll_srcStressCoeff": 0., "gamma_ll_srcHCoeff": 0.}

    def __init__(self, solver, args=None):
        super().__init__(solver, args)
        for k in self.guiVars:
            if k in self.args:
                self.vars[k] = self.args[k]
        if self.args.get("useShift", "none") != "none":
            raise NotImplementedError("adm3d: only useShift='none' is in the hot-path scope")
        for k in ("K_ll_srcStressCoeff", "K_ll_srcHCoeff", "gamma_ll_srcStressCoeff", "gamma_ll_srcHCoeff"):
            if self.vars[k] != 0:
                raise NotImplementedError("adm3d: %s != 0 is not built (reference default 0, adm3d.lua:207-216)" % k)

    def eqnParams(self):
        v = self.vars
        return [float(F_EQN.index(v["f_eqn"])), v["a_convCoeff"], v["d_convCoeff"], v["V_convCoeff"]]

    def consArray(self, W):
        """applyInitCondCell (adm3d.cl:161-194): alpha, gamma_ll, K_ll from the initial condition, everything else zero
        (a_l, d_lll, V_l are then set by initDerivs)."""
        shape = np.shape(W["alpha"])
        U = np.zeros(shape + (self.numStates,), dtype=np.float64)
        U[..., 0] = W["alpha"]
        for k, s in enumerate(SYM):
            U[..., 1 + k] = W["gamma_ll_" + s]
            U[..., 28 + k] = W["K_ll_" + s]
        U[..., 37] = W.get("rho", 0.)
        return U
