"""Equation base class: the host-side plug-in surface of hydro/eqn/eqn.lua the hot path needs.

Reference: hydro/eqn/eqn.lua:130-379 (numStates/numIntStates/numWaves, consVars, guiVars),
:464-480 (unit guiVars meter/second/kilogram/coulomb/kelvin).  In the reference the device functions
are emitted from .cl templates; here each equation's device functions are the hand-written sm_100a
headers csrc/eqn_<name>.cuh selected by ``eqnId``.
"""
import numpy as np


class Equation:
    name = None
    eqnId = None
    numStates = None
    numIntStates = None
    numWaves = None
    consVars = ()           # flattened scalar field names, in cons_t order
    roeUseFluxFromCons = True   # eqn.lua:46
    guiVars = {}

    def __init__(self, solver, args=None):
        self.solver = solver
        self.args = dict(args or {})
        # eqn.lua:468-474 units, then per-eqn guiVars
        self.vars = dict(meter=1., second=1., kilogram=1., coulomb=1., kelvin=1.)
        self.vars.update(self.guiVars)

    def applySolverVars(self, solverVars):
        """initCond.solverVars override eqn guiVars (solverbase.lua:1530-1536)."""
        for k, v in solverVars.items():
            self.vars[k] = v

    def eqnParams(self):
        """-> list of <=16 doubles laid out as csrc expects (hb_fv_desc.eqn_params)."""
        raise NotImplementedError

    def consFromPrim(self, W):
        raise NotImplementedError

    def consArray(self, W):
        """dict of prim fields -> AoS array [..., numStates] (double)."""
        U = self.consFromPrim(W)
        return np.stack([np.asarray(U[k], dtype=np.float64) + 0 * W["rho"] for k in self.consVars], axis=-1)
