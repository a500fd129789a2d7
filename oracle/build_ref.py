#!/usr/bin/env python
"""oracle/build_ref.py -- TEST INFRASTRUCTURE.  Builds oracle/_ref/ from the reference's OWN source lines, where they lie.

The reference (LuaJIT + templated OpenCL-C) cannot be built as a whole, but one piece of the hot path is
plain C inside the template: the machine-generated ADM Bona-Masso source-term block of hydro/eqn/adm3d.cl (between the
"BEGIN CUT ... adm_noZeroRows.html" / "END CUT" markers of its addSource kernel).  This script extracts those lines at build time
(nothing is copied into the repository), wraps them in a C++ function with the struct field names the block uses, and compiles
oracle/_ref/libadm3d_ref_source.so.  tests/test_adm3d.py pins the oracle's (and the CUDA kernels') tensor-form source term
against it.  oracle/_ref/ is git-ignored and travels to the GPU box with the snapshot; /root/reference is only read here.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HYDRO_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

PRELUDE = r"""
#include <cmath>
typedef double real;
struct real3 { real x, y, z; };
struct real3s3 { real xx, xy, xz, yy, yz, zz; };
struct real3x3s3 { real3s3 x, y, z; };
struct cons_t { real alpha; real3s3 gamma_ll; real3 a_l; real3x3s3 d_lll; real3s3 K_ll; real3 V_l; };
extern "C" void adm3d_ref_source(const double* U37, const double* gamma_uu6, const double* fvals /* f, f_alpha, f_alphaSq, dalpha_f, alphaSq_dalpha_f */,
	const double* matter /* rho, S, S_ll[6] */, double* deriv37)
{
	cons_t Ustate, dstate;
	{ double* p = reinterpret_cast<double*>(&Ustate); for (int i = 0; i < 37; ++i) p[i] = U37[i]; }
	{ double* p = reinterpret_cast<double*>(&dstate); for (int i = 0; i < 37; ++i) p[i] = 0; }
	cons_t* const deriv = &dstate;
	real const alpha = Ustate.alpha;
	real3s3 const gamma_ll = Ustate.gamma_ll;
	real3 const a_l = Ustate.a_l;
	real3x3s3 const d_lll = Ustate.d_lll;
	real3s3 const K_ll = Ustate.K_ll;
	real3s3 const gamma_uu = {gamma_uu6[0], gamma_uu6[1], gamma_uu6[2], gamma_uu6[3], gamma_uu6[4], gamma_uu6[5]};
	real const f = fvals[0], f_alpha = fvals[1], f_alphaSq = fvals[2], dalpha_f = fvals[3], alphaSq_dalpha_f = fvals[4];
	real const rho = matter[0], S = matter[1];
	real3s3 const S_ll = {matter[2], matter[3], matter[4], matter[5], matter[6], matter[7]};
	(void)f; (void)dalpha_f; (void)gamma_ll; (void)rho; (void)S; (void)S_ll;
"""
EPILOGUE = r"""
	{ const double* p = reinterpret_cast<const double*>(&dstate); for (int i = 0; i < 37; ++i) deriv37[i] = p[i]; }
}
"""


def main():
    src = os.path.join(REF, "hydro", "eqn", "adm3d.cl")
    if not os.path.exists(src):
        print("build_ref: %s not present (GPU box): using the prebuilt oracle/_ref if any" % src)
        return 0
    lines = open(src).read().split("\n")
    begin = [i for i, l in enumerate(lines) if "BEGIN CUT from numerical-relativity-codegen/flux_matrix_output/adm_noZeroRows.html" in l and l.strip().startswith("//")]
    # the addSource block is the BEGIN CUT that follows the "#else	//code-generated" line
    gen = [i for i, l in enumerate(lines) if l.startswith("#else") and "code-generated" in l]
    assert gen, "marker '#else //code-generated' not found"
    b = min(i for i in begin if i > gen[0])
    e = min(i for i, l in enumerate(lines) if i > b and l.strip() == "// END CUT")
    body = "\n".join(lines[b + 1:e])
    os.makedirs(OUT, exist_ok=True)
    cpp = os.path.join(OUT, "adm3d_ref_source.cpp")
    with open(cpp, "w") as f:
        f.write("// GENERATED at build time by oracle/build_ref.py from %s lines %d-%d; not tracked by git.\n" % (src, b + 2, e))
        f.write(PRELUDE + body + EPILOGUE)
    so = os.path.join(OUT, "libadm3d_ref_source.so")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-Wno-unused-variable", "-o", so, cpp])
    print("build_ref: built", so)
    return 0


if __name__ == "__main__":
    sys.exit(main())
