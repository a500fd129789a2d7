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


class HLL(Flux):
    """hydro/flux/hll.lua:5-12: name 'hll', hllCalcWaveMethod = 'Davis direct bounded' (the reference's live setting)."""
    name = "hll"
    fluxId = 1
    hllCalcWaveMethod = "Davis direct bounded"

    def __init__(self, solver, args=None):
        super().__init__(solver, args)
        m = self.args.get("hllCalcWaveMethod", self.hllCalcWaveMethod)
        if m != self.hllCalcWaveMethod:
            raise NotImplementedError("hllCalcWaveMethod %r: only 'Davis direct bounded' (hll.lua:10) is built" % (m,))


class Rusanov(Flux):
    """hydro/flux/rusanov.lua: name 'rusanov' (local Lax-Friedrichs with the cell wave speeds)."""
    name = "rusanov"
    fluxId = 2


class EulerHLLC(HLL):
    """hydro/flux/euler-hllc.lua: name 'euler-hllc', args.hllcMethod 0 | 1 | 2 (default 2), Euler only (euler-hllc.cl:7-11)."""
    name = "euler-hllc"
    fluxId = 3

    def __init__(self, solver, args=None):
        super().__init__(solver, args)
        if getattr(solver.eqn, "name", None) != "euler":
            raise ValueError("euler-hllc only works with euler eqn (euler-hllc.cl:8-10)")
        self.hllcMethod = int(self.args.get("hllcMethod", 2))
        if self.hllcMethod not in (0, 1, 2):
            raise ValueError("hllcMethod must be 0, 1 or 2")
        self.fluxParam = self.hllcMethod


fluxes = {"roe": Roe, "hll": HLL, "rusanov": Rusanov, "euler-hllc": EulerHLLC}
