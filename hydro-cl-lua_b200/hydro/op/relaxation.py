"""Relaxation (hydro/op/relaxation.lua): Jacobi sweeps of a Poisson problem on one field of UBuf."""


class Relaxation:
    name = "relaxation"
    kind = 0
    potentialField = "ePot"          # relaxation.lua:22
    stopOnEpsilon = True             # :24
    stopEpsilon = 1e-10              # :25
    maxIters = 20                    # :26 (cmdline.selfGravPoissonMaxIter or 20)
    param = 0.

    def __init__(self, solver, **args):
        self.solver = solver
        for k in ("maxIters", "stopOnEpsilon", "stopEpsilon", "potentialField"):
            if k in args:
                setattr(self, k, args[k])
        self.index = None

    def register(self, backend):
        self.index = backend.add_op(self.kind, int(self.maxIters), bool(self.stopOnEpsilon), float(self.stopEpsilon), float(self.param))

    @property
    def lastIter(self):
        return self.solver.backend.op_info(self.index)[0]

    @property
    def lastResidual(self):
        return self.solver.backend.op_info(self.index)[1]
