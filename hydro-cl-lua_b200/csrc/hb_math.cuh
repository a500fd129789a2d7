// hb_math.cuh -- scalar helpers shared by the per-equation device functions.
//
// Everything here is `HB_HD`: it compiles as __host__ __device__ under nvcc and as plain inline C++ under
// g++ (tests/host_check builds the same headers with g++ -ffp-contract=off so the per-interface / per-cell
// arithmetic can be compared with the CPU oracle without a GPU).
//
// Reference semantics being matched (file:line into the reference tree):
//   hydro/code/math.cl:175-181,221   real3_dot = a.x*b.x + (a.y*b.y + a.z*b.z)  (right-nested)
//   hydro/coord/coord.lua:720-727    coordLenSq (identity metric), summed left to right
//   OpenCL `min`/`max` on reals are the comparison forms below (no NaN quieting)
#pragma once

#if defined(__CUDACC__)
#define HB_HD __host__ __device__ __forceinline__
#define HB_D __device__ __forceinline__
#else
#define HB_HD inline
#define HB_D inline
#endif

#include "hb_rtc_compat.h"

namespace hb {

template<class real> HB_HD real rmax(real a, real b) { return a < b ? b : a; }
template<class real> HB_HD real rmin(real a, real b) { return b < a ? b : a; }

HB_HD double rsqrt_ieee(double x) { return sqrt(x); }
HB_HD float rsqrt_ieee(float x) { return sqrtf(x); }
HB_HD double rabs(double x) { return fabs(x); }
HB_HD float rabs(float x) { return fabsf(x); }

// ---- branch-free reciprocal and (reciprocal) square root for arguments known to be positive and normal.
// The compiler's IEEE 1/x and sqrt(x) carry a special-case slow path behind a branch per call; those branches cut the flux
// routine into basic blocks and serialise its long dependent chains (MUFU seed -> Newton steps).  These forms are straight-line:
// hardware seed (relative error e <= 2^-22) + ONE third-order step (error ~ e^3 = 2^-66), accurate to the last ulp or two (not
// correctly rounded).  Third order instead of two Newton steps: 3 FP64 instructions for the reciprocal instead of 4, 6 for the
// square-root pair instead of 9 -- the FP64 pipe issues one warp instruction per two cycles, and a flux has 4 to 12 of these.
HB_HD double fastRcp(double x) {
#if defined(__CUDA_ARCH__)
	double r;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
	double const e = fma(-x, r, 1.);          // 1/x = r / (1 - e) = r (1 + e + e^2 + O(e^3))
	return fma(r, fma(e, e, e), r);
#else
	return 1. / x;
#endif
}
HB_HD float fastRcp(float x) {
#if defined(__CUDA_ARCH__)
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return fmaf(r, fmaf(-x, r, 1.f), r);
#else
	return 1.f / x;
#endif
}
// y = 1/sqrt(x), s = sqrt(x) from one seed and one third-order step
HB_HD void fastRsqrt(double x, double& y, double& s) {
#if defined(__CUDA_ARCH__)
	double y0;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
	double const t = x * y0;                  // ~ sqrt(x)
	double const e = fma(-t, y0, 1.);         // x y0^2 = 1 - e;  x^-1/2 = y0 (1 - e)^-1/2 = y0 (1 + e/2 + 3 e^2/8 + O(e^3))
	double const q = e * fma(.375, e, .5);
	y = fma(y0, q, y0);
	s = fma(t, q, t);
#else
	s = sqrt(x); y = 1. / s;
#endif
}
HB_HD void fastRsqrt(float x, float& y, float& s) {
#if defined(__CUDA_ARCH__)
	float y0;
	asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(x));
	float g = x * y0, h = .5f * y0;
	float r = fmaf(-g, h, .5f);
	g = fmaf(g, r, g); h = fmaf(h, r, h);
	s = g; y = h + h;
#else
	s = sqrtf(x); y = 1.f / s;
#endif
}

// sqrt(x) for x >= 0 that may be exactly zero (s = 0 then; the reciprocal y is garbage there and the caller selects around it)
HB_HD void fastRsqrt0(double x, double& y, double& s) { fastRsqrt(x > 1e-300 ? x : 1e-300, y, s); if (!(x > 1e-300)) s = x > 0. ? sqrt(x) : 0.; }
HB_HD void fastRsqrt0(float x, float& y, float& s) { fastRsqrt(x > 1e-37f ? x : 1e-37f, y, s); if (!(x > 1e-37f)) s = x > 0.f ? sqrtf(x) : 0.f; }

template<class real> HB_HD real lenSq3(real x, real y, real z) { return x * x + y * y + z * z; }
// math.cl real3_dot: right-nested sum
template<class real> HB_HD real dot3(real ax, real ay, real az, real bx, real by, real bz) {
	return ax * bx + (ay * by + az * bz);
}

template<class real> struct inf_of;
template<> struct inf_of<double> { static HB_HD double v() { return HUGE_VAL; } };
template<> struct inf_of<float> { static HB_HD float v() { return HUGE_VALF; } };

// hydro/app.lua:614-635 limiter table, 0-based (0 = 'donor cell'); literal expressions of the table.
template<class real> HB_HD real limiter(int id, real r) {
	switch (id) {
	case 0: return real(0.);
	case 1: return real(1.);
	case 2: return r;
	case 3: return real(.5) * (real(1.) + r);
	case 4: return rmax<real>(real(0.), r) * (real(3.) * r + real(1.)) / ((r + real(1.)) * (r + real(1.)));
	case 5: return rmax<real>(real(0.), real(1.5) * (r + rabs(r)) / (r + real(2.)));
	case 6: return rmax<real>(real(0.), real(2.) * (r + rabs(r)) / (r + real(3.)));
	case 7: return rmax<real>(real(0.), rmin<real>(real(2.) * r, rmin<real>((real(1.) + real(2.) * r) / real(3.), real(2.))));
	case 8: return rmax<real>(real(0.), rmin<real>(r, real(1.)));
	case 9: return rmax<real>(real(0.), rmin<real>(r, real(1.5)));
	case 10: return real(.5) * (r * r + r) / (r * r + r + real(1.));
	case 11: return rmax<real>(real(0.), rmin<real>(real(2.) * r, rmin<real>(real(.25) + real(.75) * r, real(4.))));
	case 12: return rmax<real>(real(0.), rmax<real>(rmin<real>(real(1.5) * r, real(1.)), rmin<real>(r, real(1.5))));
	case 13: return rmax<real>(real(0.), rmin<real>(rmin<real>(real(2.) * r, real(.75) + real(.25) * r), rmin<real>(real(.25) + real(.75) * r, real(2.))));
	case 14: return (r * r + r) / (r * r + real(1.));
	case 15: return real(2.) * r / (r * r + real(1.));
	case 16: return rmax<real>(real(0.), r) * real(2.) / (real(1.) + r);
	case 17: return rmax<real>(real(0.), rmin<real>(real(2.), rmin<real>(real(.5) * (real(1.) + r), real(2.) * r)));
	case 18: return rmax<real>(real(0.), rmax<real>(rmin<real>(real(1.), real(2.) * r), rmin<real>(real(2.), r)));
	case 19: return real(.5) * (r + real(1.)) * rmin<real>(real(1.), rmin<real>(real(4.) * r / (r + real(1.)), real(4.) / (r + real(1.))));
	}
	return real(0.);
}

// Cartesian normal for side SIDE (hydro/coord/coord.lua:2357-2410): normal_t = {int side}.
//   normal_vecDotNs(n, v)  = (v[s], v[s+1], v[s+2])           -> rot<SIDE>::fwd
//   normal_vecFromNs(n, v) = (v[(3-s)%3], v[(3-s+1)%3], ...)  -> rot<SIDE>::inv
//   normal_l<j><x_i>(n)    = (s == (i-j) mod 3), i,j in 1..3  -> rot<SIDE>::sel(j, i)
template<int SIDE> struct rot {
	template<class real> static HB_HD void fwd(real const (&v)[3], real (&o)[3]) {
		o[0] = v[SIDE]; o[1] = v[(SIDE + 1) % 3]; o[2] = v[(SIDE + 2) % 3];
	}
	template<class real> static HB_HD void inv(real const (&v)[3], real (&o)[3]) {
		o[0] = v[(3 - SIDE) % 3]; o[1] = v[(3 - SIDE + 1) % 3]; o[2] = v[(3 - SIDE + 2) % 3];
	}
	static constexpr bool sel(int j, int i) { return SIDE == (((i - j) % 3) + 3) % 3; }
};

}   // namespace hb
