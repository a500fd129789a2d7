"""Roe flux plug-in: host side of hydro/flux/roe.lua (the device side is csrc/roe_flux.cuh).

Reference: hydro/flux/roe.lua:5-19 (name 'roe', usesFluxLimiter = true, entropy fix off by default),
hydro/flux/flux.lua:14-37.
"""


class Flux:
    name = None
    usesFluxLimiter = False

    def __init__(self, solver, args=None):
        self.solver = solver
        self.args = dict(args or {})


class Roe(Flux):
    name = "roe"
    fluxId = 0
    usesFluxLimiter = True
    useEntropyFluxFix = False

    def __init__(self, solver, args=None):
        super().__init__(solver, args)
        if self.args.get("useEntropyFluxFix"):
            raise NotImplementedError("Harten entropy fix (roe.cl:100-106) is off in every BASELINE config")


fluxes = {"roe": Roe}
