"""hydro/op: the operators a solver runs either side of the finite-volume step (solver.ops).

Host mirror of hydro/op/relaxation.lua (Jacobi Poisson relaxation: maxIters 20, stopOnEpsilon, stopEpsilon 1e-10, :24-26),
hydro/op/selfgrav.lua (self-gravity of euler / mhd; potential = ePot; enabled by solver.useGravity, :22,112-121) and
hydro/op/nodiv.lua with the Jacobi parent (`noDivPoissonSolver=jacobi`; mhd in more than one dimension, mhd.lua:113-119).
The reference's default NoDiv parent, poisson_krylov, lives in its un-vendored 'solver' library and is not built.
The device work is the backend's (hb_fv_add_op / hb_fv_ops_reset; kernels in csrc/hb_ops_kernels.cuh)."""
from .relaxation import Relaxation
from .selfgrav import SelfGrav
from .nodiv import NoDiv

__all__ = ["Relaxation", "SelfGrav", "NoDiv"]
