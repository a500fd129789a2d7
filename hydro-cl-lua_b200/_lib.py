"""ctypes binding of the C ABI declared in include/hydrob200.h (csrc/libhydrob200.so).

These are exactly the entry points a LuaJIT ``ffi.cdef`` binding uses; the Python host mirror
goes through the same calls, so the parity tests exercise the drop-in boundary itself.  There is no CPU
fallback: a missing library raises at load time and a missing GPU raises on the first device call.
"""
import ctypes as C
import os
import subprocess

_here = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_here, "csrc")
LIB_PATH = os.path.join(CSRC, "libhydrob200.so")

HB_OK, HB_ERR_INVALID, HB_ERR_NO_DEVICE, HB_ERR_CUDA, HB_ERR_COMPILE = 0, 1, 2, 3, 4
HB_EQN_EULER, HB_EQN_MHD, HB_EQN_ADM3D = 0, 1, 2
HB_REDUCE_MIN, HB_REDUCE_MAX, HB_REDUCE_SUM = 0, 1, 2


class HydroB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libhydrob200 error %d: %s" % (code, msg))
        self.code = code


class hb_op_desc(C.Structure):
    _fields_ = [("kind", C.c_int), ("max_iters", C.c_int), ("stop_on_epsilon", C.c_int), ("stop_epsilon", C.c_double), ("param", C.c_double)]


class hb_fv_desc(C.Structure):
    _fields_ = [
        ("eqn", C.c_int), ("dim", C.c_int), ("n", C.c_int * 3), ("global_n", C.c_int * 3),
        ("use_plm", C.c_int), ("slope_limiter", C.c_int), ("flux_limiter", C.c_int),
        ("bc", C.c_int * 6), ("rk_order", C.c_int),
        ("alphas", C.c_double * 16), ("betas", C.c_double * 16),
        ("mins", C.c_double * 3), ("maxs", C.c_double * 3),
        ("cfl", C.c_double), ("fixed_dt", C.c_double), ("use_fixed_dt", C.c_int),
        ("eqn_params", C.c_double * 16),
        ("strict_fp", C.c_int), ("use_graph", C.c_int), ("stage_kernel", C.c_int), ("flux", C.c_int), ("flux_param", C.c_int),
        ("use_ctu", C.c_int),
    ]


size3 = C.c_size_t * 3
P = C.c_void_p

# name -> (restype, argtypes); every name here must be declared in include/hydrob200.h (tests/test_abi.py checks)
SIGNATURES = {
    "hb_last_error": (C.c_char_p, []),
    "hb_version": (C.c_int, []),
    "hb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "hb_ctx_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(P)]),
    "hb_ctx_destroy": (C.c_int, [P]),
    "hb_ctx_real_bytes": (C.c_int, [P]),
    "hb_device_name": (C.c_int, [P, C.c_char_p, C.c_size_t]),
    "hb_device_max_threads": (C.c_int, [P, C.POINTER(C.c_int)]),
    "hb_device_sm_count": (C.c_int, [P, C.POINTER(C.c_int)]),
    "hb_sync": (C.c_int, [P]),
    "hb_ctx_stream": (P, [P]),
    "hb_timer_start": (C.c_int, [P]),
    "hb_timer_stop": (C.c_int, [P, C.POINTER(C.c_float)]),
    "hb_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(P)]),
    "hb_host_free": (C.c_int, [P]),
    "hb_buf_alloc": (C.c_int, [P, C.c_size_t, C.POINTER(P)]),
    "hb_buf_free": (C.c_int, [P]),
    "hb_buf_size": (C.c_size_t, [P]),
    "hb_buf_devptr": (P, [P]),
    "hb_buf_write": (C.c_int, [P, P, C.c_size_t, C.c_size_t]),
    "hb_buf_read": (C.c_int, [P, P, C.c_size_t, C.c_size_t]),
    "hb_buf_fill": (C.c_int, [P, P, C.c_size_t, C.c_size_t, C.c_size_t]),
    "hb_buf_copy": (C.c_int, [P, C.c_size_t, P, C.c_size_t, C.c_size_t]),
    "hb_buf_copy_rect": (C.c_int, [P, P, size3, size3, size3, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t]),
    "hb_module_compile": (C.c_int, [P, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(P), C.c_char_p, C.c_size_t]),
    "hb_module_free": (C.c_int, [P]),
    "hb_kernel_get": (C.c_int, [P, C.c_char_p, C.POINTER(P)]),
    "hb_kernel_set_arg": (C.c_int, [P, C.c_int, P, C.c_size_t]),
    "hb_kernel_set_arg_buf": (C.c_int, [P, C.c_int, P]),
    "hb_kernel_launch": (C.c_int, [P, size3, size3, C.c_size_t]),
    "hb_kernel_free": (C.c_int, [P]),
    "hb_cl_prelude": (C.c_char_p, []),
    "hb_cl_translate": (C.c_int, [C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "hb_module_compile_opencl": (C.c_int, [P, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(P), C.c_char_p, C.c_size_t]),
    "hb_reduce": (C.c_int, [P, P, C.c_size_t, C.c_int, C.POINTER(C.c_double)]),
    "hb_sizeof_fv_desc": (C.c_size_t, []),
    "hb_fv_create": (C.c_int, [P, C.POINTER(hb_fv_desc), C.POINTER(P)]),
    "hb_fv_destroy": (C.c_int, [P]),
    "hb_fv_create_from_source": (C.c_int, [P, C.POINTER(hb_fv_desc), C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(P), C.c_char_p, C.c_size_t]),
    "hb_fv_num_states": (C.c_int, [P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "hb_fv_num_cells": (C.c_longlong, [P]),
    "hb_fv_set_state": (C.c_int, [P, P]),
    "hb_fv_get_state": (C.c_int, [P, P]),
    "hb_fv_set_state_async": (C.c_int, [P, P]),
    "hb_fv_get_state_async": (C.c_int, [P, P]),
    "hb_fv_wait_transfers": (C.c_int, [P]),
    "hb_fv_state_devptr": (C.c_int, [P, C.POINTER(P), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "hb_fv_boundary": (C.c_int, [P]),
    "hb_sizeof_op_desc": (C.c_size_t, []),
    "hb_fv_add_op": (C.c_int, [P, C.POINTER(hb_op_desc), C.POINTER(C.c_int)]),
    "hb_fv_ops_reset": (C.c_int, [P]),
    "hb_fv_op_info": (C.c_int, [P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "hb_fv_set_fixed_boundary": (C.c_int, [P, C.c_int, C.POINTER(C.c_double), C.c_int]),
    "hb_fv_constrainU": (C.c_int, [P]),
    "hb_fv_init_derivs": (C.c_int, [P]),
    "hb_fv_calc_dt": (C.c_int, [P, C.POINTER(C.c_double)]),
    "hb_fv_step": (C.c_int, [P, C.c_double]),
    "hb_fv_update": (C.c_int, [P, C.c_int]),
    "hb_fv_get_time": (C.c_int, [P, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "hb_fv_set_time": (C.c_int, [P, C.c_double]),
    "hb_fv_calc_deriv": (C.c_int, [P, C.c_double, P]),
    "hb_fv_launch_count": (C.c_int, [P, C.POINTER(C.c_longlong)]),
    "hb_fv_describe": (C.c_int, [P, C.c_char_p, C.c_size_t]),
    "hb_rk_plan": (C.c_int, [C.c_int, P, P, C.c_int, C.c_char_p, C.c_size_t]),
    "hb_fv_profile": (C.c_int, [P, C.c_int]),
    "hb_fv_profile_read": (C.c_int, [P, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "hb_debug_eval": (C.c_int, [P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, P, P, P, C.c_size_t, P, C.c_size_t]),
    "hb_ghost_source": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "hb_comm_unique_id": (C.c_int, [C.c_char_p]),
    "hb_fv_comm_init": (C.c_int, [P, C.c_int, C.c_int, C.c_char_p]),
    "hb_fv_comm_destroy": (C.c_int, [P]),
}

_lib = None


def build(force=False, jobs=8):
    """Compile csrc/ for sm_100a with nvcc (in-tree; csrc/Makefile)."""
    if force:
        subprocess.check_call(["make", "-C", CSRC, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", CSRC, "-j%d" % jobs], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    """The loaded C-ABI library.  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HydroB200Error(HB_ERR_INVALID, "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                 "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(code):
    if code != HB_OK:
        raise HydroB200Error(code, (lib().hb_last_error() or b"").decode("utf-8", "replace"))
