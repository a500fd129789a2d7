"""The C-ABI boundary without a GPU: libhydrob200.so loads, exports every symbol include/hydrob200.h declares, the ctypes
mirror of hb_fv_desc has the C layout, and device entry points fail loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "hydrob200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-zA-Z0-9_]+)\s*\(", src)))


def test_header_declares_what_the_binding_uses(hydrob200):
    from importlib import import_module
    hb = import_module("hydro-cl-lua_b200._lib")
    decl = set(declared_functions())
    assert decl, "no declarations parsed"
    missing = [n for n in hb.SIGNATURES if n not in decl]
    assert not missing, "bound but not declared in include/hydrob200.h: %s" % missing
    unbound = [n for n in decl if n not in hb.SIGNATURES]
    assert not unbound, "declared but not bound by _lib.py: %s" % unbound


def test_library_exports_every_declared_symbol(hydrob200):
    from importlib import import_module
    hb = import_module("hydro-cl-lua_b200._lib")
    L = hb.lib()
    for name in declared_functions():
        assert hasattr(L, name), "libhydrob200.so does not export %s" % name
    assert L.hb_version() >= 1


def test_desc_layout_matches_c(hydrob200):
    from importlib import import_module
    hb = import_module("hydro-cl-lua_b200._lib")
    assert hb.lib().hb_sizeof_fv_desc() == C.sizeof(hb.hb_fv_desc)
    assert hb.lib().hb_sizeof_op_desc() == C.sizeof(hb.hb_op_desc)


def test_no_device_means_error_not_fallback(hydrob200):
    from importlib import import_module
    hb = import_module("hydro-cl-lua_b200._lib")
    L = hb.lib()
    n = C.c_int(-1)
    L.hb_device_count(C.byref(n))
    if n.value > 0:
        pytest.skip("a CUDA device is present")
    h = hb.P()
    rc = L.hb_ctx_create(0, 8, C.byref(h))
    assert rc == hb.HB_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.hb_last_error()
    with pytest.raises(hb.HydroB200Error):
        hydrob200.FiniteVolumeSolver(dict(eqn="euler", dim=1, gridSize=[16], initCond="Sod"))


def test_lua_binding_names_exist_in_header():
    """Every lib.hb_* the LuaJIT binding calls must be declared (the Lua files cannot be executed here)."""
    decl = set(declared_functions())
    hdr = open(os.path.join(ROOT, "include", "hydrob200.h")).read()
    for fn in ("ffi.lua", "env.lua", "fvsolver.lua"):
        src = open(os.path.join(ROOT, "lua", "hydrob200", fn)).read()
        for name in set(re.findall(r"lib\.(hb_[a-zA-Z0-9_]+)", src)):
            assert name in decl, "%s uses undeclared %s" % (fn, name)
        for name in set(re.findall(r"lib\.(HB_[A-Z0-9_]+)", src)):
            assert re.search(r"#define\s+%s\b" % name, hdr), "%s uses undefined constant %s" % (fn, name)
