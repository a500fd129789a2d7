"""NoDiv (hydro/op/nodiv.lua) over the Jacobi relaxation: del^2 psi = div B, then B -= grad psi, once per step after the
integrator (NoDiv:step, :180-184; SolverBase:step runs boundary() and constrainU() before it, solverbase.lua:3230-3237)."""
from .relaxation import Relaxation


class NoDiv(Relaxation):
    name = "NoDiv"
    kind = 2                          # HB_OP_NODIV
    vectorField = "B"                 # nodiv.lua:19
    potentialField = "psi"            # nodiv.lua:20
