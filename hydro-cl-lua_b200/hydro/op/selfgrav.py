"""SelfGrav (hydro/op/selfgrav.lua, selfgrav.cl): del^2 ePot = 4 pi G rho by Jacobi relaxation inside every stage's addSource,
deriv.m -= rho grad ePot, deriv.ETotal -= m . grad ePot, then ePot -= max(ePot) (offsetPotential, :123-147)."""
from .relaxation import Relaxation


class SelfGrav(Relaxation):
    name = "selfgrav"
    kind = 1                          # HB_OP_SELFGRAV
    enableField = "useGravity"        # selfgrav.lua:22
    gravitationalConstant = 1.        # selfgrav.lua:39-41, in units where unit_m3_per_kg_s2 = 1

    def __init__(self, solver, **args):
        super().__init__(solver, **args)
        self.gravitationalConstant = float(args.get("gravitationalConstant", self.gravitationalConstant))
        self.param = self.gravitationalConstant
