from .adm3d import ADM3D
from .euler import Euler
from .mhd import MHD

eqns = {"euler": Euler, "mhd": MHD, "adm3d": ADM3D}
