"""hb_cl_translate (csrc/hb_cl.cu: the OpenCL-C dialect -> CUDA rewrite behind hb_module_compile_opencl) without a GPU: the translated
source of the kernels tests/test_gpu_fine_grained.py runs on the device is compiled as C++17 by g++ under a ten-line shim (one CUDA
thread = one call, blockIdx / threadIdx as globals) and EXECUTED on the CPU against numpy.  Proves the rewrite itself -- `kernel`,
`global`, `constant`, C99 compound literals (also inside #define bodies), int4 -- is valid C++ with the dialect's meaning; NVRTC and the
launch path are the GPU test's business."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

SHIM = r'''
#include <cmath>
#include <cstring>
typedef double real;
struct int4 { int x, y, z, w; };
struct dim3_ { unsigned x, y, z; };
static thread_local dim3_ blockIdx, threadIdx, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
#define __global__
#define __device__
#define __host__
#define __constant__ static
#define __forceinline__ inline
#define get_global_id(i) ((int)((i) == 0 ? blockIdx.x * blockDim.x + threadIdx.x : (i) == 1 ? blockIdx.y * blockDim.y + threadIdx.y : blockIdx.z * blockDim.z + threadIdx.z))
'''
DRIVER = r'''
extern "C" void run_multAdd(const solver_t* s, cons_t* a, const cons_t* b, const cons_t* c, double d, int nx, int ny, int nz) {
	for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) { blockIdx = {unsigned(i), unsigned(j), unsigned(k)}; multAdd(s, a, b, c, d); }
}
extern "C" void run_multAddInto(const solver_t* s, cons_t* a, const cons_t* b, double c, int nx, int ny, int nz) {
	for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) { blockIdx = {unsigned(i), unsigned(j), unsigned(k)}; multAddInto(s, a, b, c); }
}
extern "C" int sizeof_solver() { return (int)sizeof(solver_t); }
'''


def translate(hydrob200, src):
    from importlib import import_module
    hb = import_module("hydro-cl-lua_b200._lib")
    L = hb.lib()
    need = C.c_size_t()
    out = C.create_string_buffer(1 << 20)
    hb.check(L.hb_cl_translate(src.encode(), out, len(out), C.byref(need)))
    assert 0 < need.value <= len(out)
    return out.value.decode()


def test_rewrite_rules(hydrob200):
    t = translate(hydrob200, "kernel void f(global real* a, constant solver_t const * const s, local real* w) { a[0] = 1; }\n"
                             "constant int numStates = 6;\n#define _real3(a,b,c) ((real3){.x=a, .y=b, .z=c})\n"
                             "static inline real3 g(real3 a) { return (real3){.x = a.x, .y = 2., .z = a.z}; }\n"
                             "real2 h() { return (real2){1., 2.}; }\n")
    assert 'extern "C" __global__ void f(' in t
    assert "global" not in t.replace("__global__", "")            # address-space qualifiers dropped ...
    assert "__constant__ const int numStates = 6;" in t           # ... and kept where they declare storage
    assert "{.x" not in t and "_hb_v.x=a" in t.replace(" ", "")    # compound literals with designators, also inside the #define body
    assert "real2{1., 2.}" in t                                   # positional compound literal -> brace initialisation


def test_translated_kernels_compile_and_run_on_the_cpu(hydrob200, tmp_path):
    import test_gpu_fine_grained as fg
    cpp = tmp_path / "k.cpp"
    cpp.write_text(SHIM + translate(hydrob200, fg.SIMPLE) + DRIVER)
    so = tmp_path / "k.so"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-Wno-unused-value", "-o", str(so), str(cpp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    K = C.CDLL(str(so))
    nx, ny, nz = 7, 5, 3

    class Solver(C.Structure):          # tests/test_gpu_fine_grained.py's solver_t
        _fields_ = [("gridSize", C.c_int * 4), ("stepsize", C.c_int * 4), ("mins", C.c_double * 3), ("maxs", C.c_double * 3), ("grid_dx", C.c_double * 3),
                    ("numGhost", C.c_int), ("dim", C.c_int), ("gamma", C.c_double), ("rhoMin", C.c_double), ("PMin", C.c_double), ("aov", C.c_double * 3)]
    assert K.sizeof_solver() == C.sizeof(Solver)
    s = Solver()
    s.gridSize[:] = [nx, ny, nz, 1]
    s.stepsize[:] = [1, nx, nx * ny, nx * ny * nz]
    s.dim = 3
    rng = np.random.default_rng(3)
    a = rng.uniform(-1, 1, (nz, ny, nx, 6)); b = rng.uniform(-1, 1, a.shape); c = rng.uniform(-1, 1, a.shape)
    d = .37
    K.run_multAdd.argtypes = [C.c_void_p] * 4 + [C.c_double] + [C.c_int] * 3
    K.run_multAddInto.argtypes = [C.c_void_p] * 3 + [C.c_double] + [C.c_int] * 3
    got = a.copy()
    K.run_multAdd(C.byref(s), got.ctypes.data, b.ctypes.data, c.ctypes.data, d, nx, ny, nz)
    assert np.array_equal(got, b + c * d)
    got = a.copy()
    K.run_multAddInto(C.byref(s), got.ctypes.data, b.ctypes.data, d, nx, ny, nz)
    assert np.array_equal(got, a + b * d)
