"""FiniteVolumeSolver: flux plug-in + calcDeriv (hydro/solver/fvsolver.lua).

Reference: fvsolver.lua:16-52 (class, createFlux), :57-198 (calcFlux kernel), :225-302 (calcDeriv =
[calcLR] -> calcFlux -> calcDerivFromFlux).  Here calcDeriv, the RK stage combination, constrainU, the
ghost fill and the CFL reduction are fused on the device (csrc/fv_stage.cuh); this class keeps the
reference's constructor arguments (config.lua:8-690) and methods.
"""
from ..flux.roe import fluxes
from .gridsolver import GridSolver


class FiniteVolumeSolver(GridSolver):
    name = "fvsolver"

    def initObjs(self, args):
        super().initObjs(args)
        fluxArgs = dict(args.get("fluxArgs") or {})
        if "hllcMethod" in args:      # tests/test-order/schemes.lua passes it at the top level of the solver args
            fluxArgs.setdefault("hllcMethod", args["hllcMethod"])
        self.createFlux(args.get("flux", "roe"), fluxArgs)
        if not self.flux.usesFluxLimiter:
            self.fluxLimiter = 0

    def createFlux(self, fluxName, fluxArgs=None):
        if fluxName not in fluxes:
            raise NotImplementedError("flux %r is not built (SURVEY 8f2: 'roe', 'hll', 'rusanov' are)" % (fluxName,))
        self.flux = fluxes[fluxName](self, fluxArgs)

    def createBackend(self, args):
        backend = args.get("backend")
        if backend is not None:
            return backend(self)        # tests inject the CPU oracle here; the product never does
        from ...backend import CudaBackend
        return CudaBackend(self, device=args.get("device", 0), comm=args.get("comm"))

    def calcDeriv(self, dt):
        """fvsolver.lua:225-302: returns dU/dt (AoS, float64) of the current state.  Debug / test entry point."""
        return self.backend.calc_deriv(dt)
