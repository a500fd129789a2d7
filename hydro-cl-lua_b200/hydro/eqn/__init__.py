from .euler import Euler
from .mhd import MHD

eqns = {"euler": Euler, "mhd": MHD}
