"""hydro-cl-lua_b200: a B200-native (sm_100a CUDA) backend for hydro-cl-lua's explicit finite-volume update.

Layout:
  csrc/      hand-written CUDA kernels + the C-ABI shim (include/hydrob200.h) -> csrc/libhydrob200.so
  _lib.py    ctypes binding of the C-ABI (the same entry points a LuaJIT FFI binding would use)
  backend.py the device backend used by the host-side mirror below
  hydro/     host-side mirror of the reference's plug-in surface (hydro/solver, hydro/eqn, hydro/flux,
             hydro/int, hydro/init): same names, argument meaning and error behaviour

There is no CPU fallback: importing works without a GPU (so CPU-only tests can check the host logic and
the C-ABI symbols), but any device operation raises if the CUDA library or a GPU is missing.
"""
from .hydro.solver.fvsolver import FiniteVolumeSolver   # noqa: F401
from .hydro import app   # noqa: F401

__all__ = ["FiniteVolumeSolver", "app"]
