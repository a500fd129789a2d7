"""Import alias: ``import hydrob200`` == the package in ``hydro-cl-lua_b200/``.

The package directory carries the reference's name (with hyphens), which the ``import`` statement
cannot spell; this module loads it through importlib and re-exports it.
"""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("hydro-cl-lua_b200")
sys.modules[__name__] = _pkg
