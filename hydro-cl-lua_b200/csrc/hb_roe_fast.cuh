// hb_roe_fast.cuh -- production ("fast") forms of the per-interface and per-cell arithmetic.
//
// The literal functions (hb_roe.cuh, hb_eqn_*.cuh) follow the reference's OpenCL-C expression by expression so that the
// -fmad=false build is bit-identical to the CPU oracle.  On the B200 that literal form is bound by instruction issue, not by
// HBM: one Euler Roe flux is ~1100 SASS instructions (4 reciprocals, 2 divisions, 3 square roots, the vacuum branches, a
// full 5x5 left and right transform), one 'plm cons' slope ~100 (a division inside a 20-way limiter switch).
// The production build (Eqn::FAST) evaluates the SAME mathematical expressions, re-associated:
//   * slope limiters that are symmetric (phi(r)/r = phi(1/r): minmod, superbee) need no ratio: sigma = f(dUL, dUR);
//   * the Roe flux of the Euler equations is written in wave-strength form (alpha = L dU in closed form, R |Lambda| alpha
//     accumulated by groups), with branch-free reciprocal / reciprocal-square-root forms (4 per interface, two of them
//     independent, instead of 4 reciprocals + 2 divisions + 3 square roots behind slow-path branches); states that take any of the reference's
//     special branches (rho < 1e-5 on either side, rho < rhoMin, h < 1e-5: hydro/eqn/euler.cl:357-426,442,501) are sent to
//     the literal code, out of line;
//   * constrainU + calcDTCell share the primitive recovery; the CFL reduction tracks max(lambda/dx) and inverts once.
// Results agree with the literal form to rounding (a few ulp per flux); the parity tests hold the production build to
// 1e-12 relative L-infinity against the oracle after N steps (BASELINE.json north_star), the strict build to bit-exactness.
#pragma once
#include "hb_roe.cuh"
#include "hb_eqn_euler.cuh"
#include "hb_eqn_mhd.cuh"
#include "hb_rtc_compat.h"

#if defined(__CUDACC__)
#define HB_NOINLINE __host__ __device__ __noinline__
#else
#define HB_NOINLINE
#endif

namespace hb {

// ---- 'plm cons' half slope (hydro/solver/plm.cl:56-76) with the limiter fixed at compile time.
// minmod (hydro/app.lua:622): phi = max(0, min(r, 1)); superbee (:632): phi = max(0, max(min(1, 2r), min(2, r))).
// Both satisfy phi(r) d = phi(1/r) n for r = n/d, so which of dUL, dUR is the denominator (the dUC >= 0 test) does not matter,
// and sigma = 0 when n d <= 0 (this also covers the exact d == 0 test).
template<class real, int LIM, bool FAST> HB_HD real plmHalfSlopeT(int lim, real UL, real U, real UR) {
	if (FAST && LIM == 8) {
		real const dR = UR - U, dL = U - UL;
		real const m = rabs(dL) < rabs(dR) ? dL : dR;
		return dL * dR > real(0) ? real(.5) * m : real(0);
	} else if (FAST && LIM == 18) {
		real const dR = UR - U, dL = U - UL;
		real const a = rabs(dL), b = rabs(dR);
		real const m = rmax<real>(rmin<real>(real(2.) * a, b), rmin<real>(a, real(2.) * b));
		real const sg = dR > real(0) ? real(.5) : real(-.5);
		return dL * dR > real(0) ? sg * m : real(0);
	} else {
		return plmHalfSlope<real>(lim, UL, U, UR);
	}
}

// Both face states of cell b from its neighbours a, c ('plm cons', plm.cl:56-76): lo = U - .5 sigma, hi = U + .5 sigma.
// Production minmod in double: sigma/2 = h m with m = the difference of smaller magnitude (ONE FP64 compare) and h = .5 when the
// two differences have the same sign bit, else 0 (integer test on the high words; a zero difference gives m = 0 either way), so a
// face state is one FMA: 4 FP64 instructions per variable (2 DADD, DSETP, 2 DFMA share) instead of 8.
HB_HD bool sameSignBit(double a, double b) {
#if defined(__CUDA_ARCH__)
	return (__double2hiint(a) ^ __double2hiint(b)) >= 0;
#else
	return std::signbit(a) == std::signbit(b);
#endif
}
// x > 0 and the high 32 bits of x above those of the (positive) threshold t: implies x >= t (doubles); plain comparison in float
HB_HD bool hiWordAbove(double x, double t) {
#if defined(__CUDA_ARCH__)
	return __double2hiint(x) > __double2hiint(t);
#else
	int64_t a, b; memcpy(&a, &x, 8); memcpy(&b, &t, 8);
	return int32_t(a >> 32) > int32_t(b >> 32);
#endif
}
HB_HD bool hiWordAbove(float x, float t) { return x >= t; }
template<class real, int LIM, bool FAST> HB_HD void plmCellFacesT(int lim, real a, real b, real c, real& lo, real& hi) {
	if constexpr (FAST && LIM == 8 && sizeof(real) == 8) {
		real const dL = b - a, dR = c - b;
		real const m = rabs(dL) < rabs(dR) ? dL : dR;
		lo = b; hi = b;
		if (sameSignBit(dL, dR)) { lo = fma(real(-.5), m, b); hi = fma(real(.5), m, b); }
	} else {
		real const s = plmHalfSlopeT<real, LIM, FAST>(lim, a, b, c);
		lo = b - s;
		hi = b + s;
	}
}
// The two face states of the interface between cells b and c from the four-cell stencil a, b, c, d: UL = b + .5 sigma(b), UR = c - .5 sigma(c).
template<class real, int LIM, bool FAST> HB_HD void plmFacesT(int lim, real a, real b, real c, real d, real& UL, real& UR) {
	if constexpr (FAST && LIM == 8 && sizeof(real) == 8) {
		real const d1 = b - a, d2 = c - b, d3 = d - c;
		real const m1 = rabs(d1) < rabs(d2) ? d1 : d2;
		real const m2 = rabs(d2) < rabs(d3) ? d2 : d3;
		UL = b; UR = c;
		if (sameSignBit(d1, d2)) UL = fma(real(.5), m1, b);
		if (sameSignBit(d2, d3)) UR = fma(real(-.5), m2, c);
	} else {
		UL = b + plmHalfSlopeT<real, LIM, FAST>(lim, a, b, c);
		UR = c - plmHalfSlopeT<real, LIM, FAST>(lim, b, c, d);
	}
}

template<class real, int n> struct VecN { real v[n]; };

// OUTLINE: call one out-of-line copy of the (large) production flux core instead of inlining it per side
template<class Eqn, int SIDE, bool OUTLINE = false>
HB_HD void roeFluxAuto(typename Eqn::real (&F)[Eqn::nI], typename Eqn::Params const& s,
	typename Eqn::real const (&UL)[Eqn::nI], typename Eqn::real const (&UR)[Eqn::nI]);

// the literal Roe flux, out of line: the rare special-branch states of the production kernels go here
template<class Eqn, int SIDE>
HB_NOINLINE VecN<typename Eqn::real, Eqn::nI> roeFluxOutOfLine(typename Eqn::Params s, VecN<typename Eqn::real, Eqn::nI> UL, VecN<typename Eqn::real, Eqn::nI> UR) {
	VecN<typename Eqn::real, Eqn::nI> F;
	roeFlux<Eqn, SIDE>(F.v, s, UL.v, UR.v);
	return F;
}

// Roe flux of the Euler equations, wave-strength form.  Same quantities as hydro/eqn/euler.cl:347-542 + hydro/flux/roe.cl:17-163
// with useFluxLimiter == false and roeUseFluxFromCons == true:
//   F = .5 (F(UL) + F(UR)) - .5 sum_j |lambda_j| alpha_j r_j,   alpha = L (UR - UL)
//   G = (gamma-1) (.5 v^2 drho - v.dm + dE),  K = Cs (dm_n - v_n drho)
//   alpha_0,4 = (G -/+ K) / (2 Cs^2),  alpha_1 = drho - G / Cs^2,  alpha_2,3 = dm_t - v_t drho
// The core is straight-line code (no branch): it returns false when the state pair takes one of the reference's special branches,
// in which case F is meaningless and the caller substitutes the literal flux (eulerRoeFluxFixup).  Two cores issued back to back
// (x and y interfaces of a cell) interleave in the instruction stream, which hides the latency of their dependent chains.
template<class Eqn, int SIDE>
HB_HD bool eulerRoeFluxCore(typename Eqn::real (&F)[5], typename Eqn::Params const& s, typename Eqn::real const (&UL)[5], typename Eqn::real const (&UR)[5])
{
	typedef typename Eqn::real real;
	constexpr int n = SIDE, t1 = (SIDE + 1) % 3, t2 = (SIDE + 2) % 3;
	real const rhoL = UL[0], rhoR = UR[0];
	real const g1 = s.gamma_1;
	// sqrt(rho) and 1/rho of both sides from two independent reciprocal square roots (1/rho = (1/sqrt(rho))^2)
	real yL, sL, yR, sR;
	fastRsqrt(rhoL, yL, sL);
	fastRsqrt(rhoR, yR, sR);
	real const iL = yL * yL, iR = yR * yR;
	real vL[3], vR[3];
	#pragma unroll
	for (int q = 0; q < 3; ++q) { vL[q] = UL[1 + q] * iL; vR[q] = UR[1 + q] * iR; }
	real const PL = g1 * (UL[4] - real(.5) * (UL[1] * vL[0] + UL[2] * vL[1] + UL[3] * vL[2]));
	real const PR = g1 * (UR[4] - real(.5) * (UR[1] * vR[0] + UR[2] * vR[1] + UR[3] * vR[2]));
	real const HL = UL[4] + PL, HR = UR[4] + PR;             // rho hTotal
	// Roe weights w = sqrt(rho) / (sqrt(rhoL) + sqrt(rhoR))
	real const iS = fastRcp(sL + sR);
	real const wL = sL * iS, wR = sR * iS;
	real v[3];
	#pragma unroll
	for (int q = 0; q < 3; ++q) v[q] = vL[q] * wL + vR[q] * wR;
	real const H = (HL * yL + HR * yR) * iS;                 // hTotalL wL + hTotalR wR: (1/rho) sqrt(rho) = 1/sqrt(rho)
	real const eK = real(.5) * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
	real const h = H - eK;
	// the reference's special branches (euler.cl:357-426: rhoEpsilon = 1e-5; :442,501: rho < rhoMin).  The test is conservative and
	// on the integer pipe: "high word above the threshold's high word" implies x >= threshold for positive x; a state pair within 2^-20
	// of a threshold, a negative or a zero operand fail it and take the literal code, which is right for every state.
	bool const regular = hiWordAbove(rhoL, s.rhoFloor) && hiWordAbove(rhoR, s.rhoFloor) && hiWordAbove(h, real(1e-5));
	real const Cs2 = g1 * h;
	real Cs, yC;                                             // Cs, 1 / Cs
	fastRsqrt(Cs2, yC, Cs);
	real const iCs2 = yC * yC;
	real const drho = rhoR - rhoL, dE = UR[4] - UL[4];
	real dm[3];
	#pragma unroll
	for (int q = 0; q < 3; ++q) dm[q] = UR[1 + q] - UL[1 + q];
	// wave strengths times |lambda|, pre-multiplied by the flux's 1/2: b_j = .5 |lambda_j| alpha_j
	real const Gq = s.quarterGamma_1 * (eK * drho - (v[0] * dm[0] + v[1] * dm[1] + v[2] * dm[2]) + dE) * iCs2;   // G / (4 Cs^2)
	real const Kq = (dm[n] - v[n] * drho) * (real(.25) * yC);                                                   // K / (4 Cs^2)
	real const b0 = rabs(v[n] - Cs) * (Gq - Kq);
	real const b4 = rabs(v[n] + Cs) * (Gq + Kq);
	real const lamh = real(.5) * rabs(v[n]);
	real const b1 = lamh * fma(real(-4.), Gq, drho);
	real const b2 = lamh * (dm[t1] - v[t1] * drho);
	real const b3 = lamh * (dm[t2] - v[t2] * drho);
	real const b04 = b0 + b4;
	real const sum = b04 + b1, dif = (b4 - b0) * Cs;
	real const vnL = vL[n], vnR = vR[n];
	F[0] = fma(real(.5), UL[1 + n] + UR[1 + n], -sum);
	F[1 + n] = fma(real(.5), (UL[1 + n] * vnL + UR[1 + n] * vnR) + (PL + PR), -fma(sum, v[n], dif));
	F[1 + t1] = fma(real(.5), UL[1 + t1] * vnL + UR[1 + t1] * vnR, -fma(sum, v[t1], b2));
	F[1 + t2] = fma(real(.5), UL[1 + t2] * vnL + UR[1 + t2] * vnR, -fma(sum, v[t2], b3));
	F[4] = fma(real(.5), HL * vnL + HR * vnR, -(b04 * H + dif * v[n] + b1 * eK + b2 * v[t1] + b3 * v[t2]));
	return regular;
}

template<class Eqn, int SIDE>
HB_HD void eulerRoeFluxFixup(bool regular, typename Eqn::real (&F)[5], typename Eqn::Params const& s, typename Eqn::real const (&UL)[5], typename Eqn::real const (&UR)[5])
{
	typedef typename Eqn::real real;
	if (!regular) {
		VecN<real, 5> a, b;
		#pragma unroll
		for (int q = 0; q < 5; ++q) { a.v[q] = UL[q]; b.v[q] = UR[q]; }
		VecN<real, 5> const r = roeFluxOutOfLine<Eqn, SIDE>(s, a, b);
		#pragma unroll
		for (int q = 0; q < 5; ++q) F[q] = r.v[q];
	}
}

template<class Eqn, int SIDE>
HB_HD void eulerRoeFluxFast(typename Eqn::real (&F)[5], typename Eqn::Params const& s, typename Eqn::real const (&UL)[5], typename Eqn::real const (&UR)[5])
{
	bool const regular = eulerRoeFluxCore<Eqn, SIDE>(F, s, UL, UR);
	eulerRoeFluxFixup<Eqn, SIDE>(regular, F, s, UL, UR);
}

// Roe flux of ideal MHD, production form.  Same quantities as hydro/eqn/mhd.cl:296-436 (calcRoeValues incl. the swapped sqrt(rho)
// weights of B.y, B.z at :335-336; eigen_forRoeAvgs), :575-783 (left / right transforms), :441-467 (fluxFromCons) and
// hydro/flux/roe.cl:17-163 with useFluxLimiter == false, roeUseFluxFromCons == true, re-associated:
//   * 10 reciprocal square roots + 2 reciprocals (branch-free seeds + Newton steps) replace 13 square roots and ~30 divisions:
//     1/rho = (1/sqrt(rhoL))(1/sqrt(rhoR)), CAx = |B.x| / sqrt(rho), Cs = aTilde CAx / Cf, BStarPerpLen = BPerpLen sqrt(gamma_1 - gamma_2 Y),
//     betaStar / betaStarSq = beta sqrt(gamma_1 - gamma_2 Y) (beta is a unit vector), 1/mu0 and gamma_2/gamma_1 precomputed;
//   * the wave pairs (fast -/+, slow -/+, Alfven -/+) share the symmetric and antisymmetric halves of their rows of L and columns of R;
//   * the degenerate cases of the reference (BPerpLen == 0; the three alphaF/alphaS tests) are kept, as selects.
template<class Eqn, int SIDE>
HB_HD void mhdRoeFluxFast(typename Eqn::real (&F)[8], typename Eqn::Params const& s, typename Eqn::real const (&UL)[8], typename Eqn::real const (&UR)[8])
{
	typedef typename Eqn::real real;
	constexpr int n = SIDE, t1 = (SIDE + 1) % 3, t2 = (SIDE + 2) % 3;
	real const g1 = s.gamma - real(1.), g2 = s.gamma - real(2.), g2_g1 = s.g2_g1, iMu0 = s.iMu0;
	// ---- primitives of both sides (mhd.cl:160-179; the floor on rho does not enter the flux)
	real yL, sL, yR, sR;
	fastRsqrt(UL[0], yL, sL);
	fastRsqrt(UR[0], yR, sR);
	real const iL = yL * yL, iR = yR * yR;
	real vL[3], vR[3];
	#pragma unroll
	for (int q = 0; q < 3; ++q) { vL[q] = UL[1 + q] * iL; vR[q] = UR[1 + q] * iR; }
	real const vSqL = lenSq3(vL[0], vL[1], vL[2]), vSqR = lenSq3(vR[0], vR[1], vR[2]);
	real const PMagL = real(.5) * lenSq3(UL[5], UL[6], UL[7]), PMagR = real(.5) * lenSq3(UR[5], UR[6], UR[7]);
	real const PL = rmax<real>((UL[4] - real(.5) * UL[0] * vSqL - PMagL * iMu0) * g1, real(1e-7));
	real const PR = rmax<real>((UR[4] - real(.5) * UR[0] * vSqR - PMagR * iMu0) * g1, real(1e-7));
	real const hTL = (UL[4] + PL + PMagL) * iL, hTR = (UR[4] + PR + PMagR) * iR;
	// ---- Roe averages (mhd.cl:296-340)
	real const iS = fastRcp(sL + sR);
	real const wL = sL * iS, wR = sR * iS;
	real const rho = sL * sR, iRho = yL * yR;
	real const vx = vL[n] * wL + vR[n] * wR, vy = vL[t1] * wL + vR[t1] * wR, vz = vL[t2] * wL + vR[t2] * wR;
	real const hTotal = hTL * wL + hTR * wR;
	real const Bx = UL[5 + n] * wL + UR[5 + n] * wR;
	real const By = UL[5 + t1] * wR + UR[5 + t1] * wL;
	real const Bz = UL[5 + t2] * wR + UR[5 + t2] * wL;
	real const dby = UL[5 + t1] - UR[5 + t1], dbz = UL[5 + t2] - UR[5 + t2];
	real const X = real(.5) * (dby * dby + dbz * dbz) * (iS * iS);
	real const Y = real(.5) * (UL[0] + UR[0]) * iRho;
	// ---- eigensystem scalars (mhd.cl:346-436)
	real const vSq = lenSq3(vx, vy, vz);
	real const BPerpSq = By * By + Bz * Bz;
	real const gY = g1 - g2 * Y;
	real const CAxSq = Bx * Bx * iRho;
	real const hHydro = hTotal - (CAxSq + BPerpSq * iRho);
	real const aTildeSq = rmax<real>(g1 * (hHydro - real(.5) * vSq) - g2 * X, real(1e-20));
	real const BSPr = gY * BPerpSq * iRho;
	real const CATildeSq = CAxSq + BSPr;
	real const dHalf = real(.5) * (CATildeSq - aTildeSq);
	real yTmp, sqrtDiscr;
	fastRsqrt0(dHalf * dHalf + aTildeSq * BSPr, yTmp, sqrtDiscr);
	real const CfSq = real(.5) * (CATildeSq + aTildeSq) + sqrtDiscr;
	real yCf, Cf, yA, aTilde, ySR, sqrtRho, yBp, BPerpLen, ySq, sq;
	fastRsqrt(CfSq, yCf, Cf);
	fastRsqrt(aTildeSq, yA, aTilde);
	fastRsqrt(rho, ySR, sqrtRho);
	fastRsqrt0(BPerpSq, yBp, BPerpLen);
	fastRsqrt(gY, ySq, sq);
	real const CAx = rabs(Bx) * ySR;
	real const CsSq = aTildeSq * CAxSq * (yCf * yCf);
	real const Cs = aTilde * CAx * yCf;
	real const BStarPerpLen = BPerpLen * sq;
	bool const noPerp = BPerpLen == real(0);
	real const betaY = noPerp ? real(1) : By * yBp, betaZ = noPerp ? real(0) : Bz * yBp;
	real const betaStarY = betaY * ySq, betaStarZ = betaZ * ySq;
	real const betaStarSq = betaStarY * betaStarY + betaStarZ * betaStarZ;
	real const QStarY = betaY * sq, QStarZ = betaZ * sq;
	real alphaF, alphaS;
	{
		real const den = CfSq - CsSq, numF = aTildeSq - CsSq, numS = CfSq - aTildeSq;
		real const iDen = fastRcp(den);
		real y0, aF, aS;
		fastRsqrt0(numF * iDen, y0, aF);
		fastRsqrt0(numS * iDen, y0, aS);
		bool const one = den == real(0) || (numF > real(0) && numS <= real(0));
		bool const zero = den != real(0) && numF <= real(0);
		alphaF = one ? real(1) : (zero ? real(0) : aF);
		alphaS = one ? real(0) : (zero ? real(1) : aS);
	}
	real const sbx = Bx >= real(0) ? real(1) : real(-1);
	real const Qf = Cf * alphaF * sbx, Qs = Cs * alphaS * sbx;
	real const Af = aTilde * alphaF * ySR, As = aTilde * alphaS * ySR;
	// ---- characteristic differences dUe = L (UR - UL) (mhd.cl:575-679), by wave pairs
	real const norm = real(.5) * (yA * yA);
	real const Cff = norm * alphaF * Cf, Css = norm * alphaS * Cs;
	real const Qf2 = Qf * norm, Qs2 = Qs * norm;
	real const AHatF = norm * Af * rho, AHatS = norm * As * rho;
	real const afpb = norm * Af * BStarPerpLen, aspb = norm * As * BStarPerpLen;
	real const norm2 = norm * g1, alphaF2 = alphaF * norm2, alphaS2 = alphaS * norm2, norm3 = norm2 * real(2.);
	real const vqstr = vy * QStarY + vz * QStarZ;
	real const drho = UR[0] - UL[0], dE = UR[4] - UL[4];
	real const dm0 = UR[1 + n] - UL[1 + n], dm1 = UR[1 + t1] - UL[1 + t1], dm2 = UR[1 + t2] - UL[1 + t2];
	real const dB1 = UR[5 + t1] - UL[5 + t1], dB2 = UR[5 + t2] - UL[5 + t2];
	real const mv = dm0 * vx + dm1 * vy + dm2 * vz;
	real const qm = dm1 * QStarY + dm2 * QStarZ;
	real const T = drho * (vSq - hHydro) - mv + dE;
	real const symF = alphaF2 * T + drho * (Cff * Cf - aspb) + dB1 * (AHatS * QStarY - alphaF2 * By) + dB2 * (AHatS * QStarZ - alphaF2 * Bz);
	real const symS = alphaS2 * T + drho * (Css * Cs + afpb) - dB1 * (AHatF * QStarY + alphaS2 * By) - dB2 * (AHatF * QStarZ + alphaS2 * Bz);
	real const antiF = drho * (Cff * vx - Qs2 * vqstr) - dm0 * Cff + Qs2 * qm;
	real const antiS = drho * (Css * vx + Qf2 * vqstr) - dm0 * Css - Qf2 * qm;
	real const A15 = real(.5) * (drho * (vy * betaZ - vz * betaY) + dm1 * (betaZ * s.l23s) + dm2 * betaY);   // l23s: hb_eqn_mhd.cuh Params
	real const B15 = real(.5) * sqrtRho * sbx * (dB2 * betaY - dB1 * betaZ);
	real const r3 = drho * (real(1.) - norm3 * (real(.5) * vSq - g2_g1 * X)) + norm3 * (mv - dE + dB1 * By + dB2 * Bz);
	// ---- wave strengths a_j = -.5 |lambda_j| dUe_j (roe.cl:91-134 without the limiter term), summed / differenced by pairs
	real const a0 = real(-.5) * rabs(vx - Cf) * (symF + antiF), a6 = real(-.5) * rabs(vx + Cf) * (symF - antiF);
	real const a2 = real(-.5) * rabs(vx - Cs) * (symS + antiS), a4 = real(-.5) * rabs(vx + Cs) * (symS - antiS);
	real const a1 = real(-.5) * rabs(vx - CAx) * (A15 + B15), a5 = real(-.5) * rabs(vx + CAx) * (B15 - A15);
	real const a3 = real(-.5) * rabs(vx) * r3;
	real const pf = a0 + a6, mf = a6 - a0, ps = a2 + a4, ms = a4 - a2, pa = a1 + a5, ma = a5 - a1;
	// ---- F = R a (mhd.cl:683-783)
	real const Frho = alphaF * pf + alphaS * ps + a3;
	real const cross = Qf * ms - Qs * mf;
	real const vDotBeta = vy * betaStarY + vz * betaStarZ;
	real const bsq = BStarPerpLen * betaStarSq;
	real const Rm0 = vx * Frho + alphaF * Cf * mf + alphaS * Cs * ms;
	real const Rm1 = vy * Frho + betaStarY * cross + betaZ * ma;
	real const Rm2 = vz * Frho + betaStarZ * cross - betaY * ma;
	real const RE = hHydro * (alphaF * pf + alphaS * ps) + bsq * (As * pf - Af * ps) + vx * (alphaF * Cf * mf + alphaS * Cs * ms)
		+ vDotBeta * cross + (vy * betaZ - vz * betaY) * ma + a3 * (real(.5) * vSq + g2_g1 * X);
	real const mag = As * pf - Af * ps, alf = sbx * ySR * pa;
	real const RB1 = betaStarY * mag - betaZ * alf;
	real const RB2 = betaStarZ * mag + betaY * alf;
	// ---- + .5 (F(UL) + F(UR)) (mhd.cl:441-467)
	real const vnL = vL[n], vnR = vR[n], BnL = UL[5 + n], BnR = UR[5 + n];
	real const PTL = PL + PMagL * iMu0, PTR = PR + PMagR * iMu0;
	real const bL = BnL * iMu0, bR = BnR * iMu0;
	real const BdVL = dot3(UL[5], UL[6], UL[7], vL[0], vL[1], vL[2]), BdVR = dot3(UR[5], UR[6], UR[7], vR[0], vR[1], vR[2]);
	F[0] = Frho + real(.5) * (UL[1 + n] + UR[1 + n]);
	F[1 + n] = Rm0 + real(.5) * ((UL[1 + n] * vnL - UL[5 + n] * bL + PTL) + (UR[1 + n] * vnR - UR[5 + n] * bR + PTR));
	F[1 + t1] = Rm1 + real(.5) * ((UL[1 + t1] * vnL - UL[5 + t1] * bL) + (UR[1 + t1] * vnR - UR[5 + t1] * bR));
	F[1 + t2] = Rm2 + real(.5) * ((UL[1 + t2] * vnL - UL[5 + t2] * bL) + (UR[1 + t2] * vnR - UR[5 + t2] * bR));
	F[4] = RE + real(.5) * (((UL[4] + PTL) * vnL - BdVL * bL) + ((UR[4] + PTR) * vnR - BdVR * bR));
	F[5 + n] = real(0);
	F[5 + t1] = RB1 + real(.5) * ((UL[5 + t1] * vnL - vL[t1] * BnL) + (UR[5 + t1] * vnR - vR[t1] * BnR));
	F[5 + t2] = RB2 + real(.5) * ((UL[5 + t2] * vnL - vL[t2] * BnL) + (UR[5 + t2] * vnR - vR[t2] * BnR));
}

template<class Eqn>
HB_NOINLINE VecN<typename Eqn::real, 8> mhdRoeFluxFastOutOfLine(typename Eqn::Params s, VecN<typename Eqn::real, 8> UL, VecN<typename Eqn::real, 8> UR) {
	VecN<typename Eqn::real, 8> F;
	mhdRoeFluxFast<Eqn, 0>(F.v, s, UL.v, UR.v);
	return F;
}

// two independent interfaces (sides SA, SB) of one cell
template<class Eqn, int SA, int SB>
HB_HD void roeFluxPairAuto(typename Eqn::real (&FA)[Eqn::nI], typename Eqn::real (&FB)[Eqn::nI], typename Eqn::Params const& s,
	typename Eqn::real const (&ULA)[Eqn::nI], typename Eqn::real const (&URA)[Eqn::nI],
	typename Eqn::real const (&ULB)[Eqn::nI], typename Eqn::real const (&URB)[Eqn::nI])
{
	if constexpr (Eqn::FAST && Eqn::eqnId == 0) {
		bool const ra = eulerRoeFluxCore<Eqn, SA>(FA, s, ULA, URA);
		bool const rb = eulerRoeFluxCore<Eqn, SB>(FB, s, ULB, URB);
		eulerRoeFluxFixup<Eqn, SA>(ra, FA, s, ULA, URA);
		eulerRoeFluxFixup<Eqn, SB>(rb, FB, s, ULB, URB);
	} else {
		roeFluxAuto<Eqn, SA>(FA, s, ULA, URA);
		roeFluxAuto<Eqn, SB>(FB, s, ULB, URB);
	}
}

// three independent interfaces (the low faces of one cell along SA, SB, SC)
template<class Eqn, int SA, int SB, int SC>
HB_HD void roeFluxTripleAuto(typename Eqn::real (&FA)[Eqn::nI], typename Eqn::real (&FB)[Eqn::nI], typename Eqn::real (&FC)[Eqn::nI],
	typename Eqn::Params const& s,
	typename Eqn::real const (&ULA)[Eqn::nI], typename Eqn::real const (&URA)[Eqn::nI],
	typename Eqn::real const (&ULB)[Eqn::nI], typename Eqn::real const (&URB)[Eqn::nI],
	typename Eqn::real const (&ULC)[Eqn::nI], typename Eqn::real const (&URC)[Eqn::nI])
{
	if constexpr (Eqn::FAST && Eqn::eqnId == 0) {
		bool const ra = eulerRoeFluxCore<Eqn, SA>(FA, s, ULA, URA);
		bool const rb = eulerRoeFluxCore<Eqn, SB>(FB, s, ULB, URB);
		bool const rc = eulerRoeFluxCore<Eqn, SC>(FC, s, ULC, URC);
		eulerRoeFluxFixup<Eqn, SA>(ra, FA, s, ULA, URA);
		eulerRoeFluxFixup<Eqn, SB>(rb, FB, s, ULB, URB);
		eulerRoeFluxFixup<Eqn, SC>(rc, FC, s, ULC, URC);
	} else {
		roeFluxAuto<Eqn, SA>(FA, s, ULA, URA);
		roeFluxAuto<Eqn, SB>(FB, s, ULB, URB);
		roeFluxAuto<Eqn, SC>(FC, s, ULC, URC);
	}
}

template<class Eqn, int SIDE, bool OUTLINE>
HB_HD void roeFluxAuto(typename Eqn::real (&F)[Eqn::nI], typename Eqn::Params const& s,
	typename Eqn::real const (&UL)[Eqn::nI], typename Eqn::real const (&UR)[Eqn::nI])
{
	if constexpr (Eqn::FAST && Eqn::eqnId == 0) eulerRoeFluxFast<Eqn, SIDE>(F, s, UL, UR);
	else if constexpr (Eqn::FAST && Eqn::eqnId == 1) {
#if defined(__CUDA_ARCH__)
		// OUTLINE: one out-of-line copy of the flux core serves every side: the operands are handed over rotated into the interface
		// frame (normal = component 0) and the result rotated back -- register renaming on either side of the call.  Inlined per side
		// three ~600-instruction copies push the 3-D plane loop past the instruction cache (measured, 256 x 256 x 64 MHD stage:
		// 1.41 ms inlined, 1.28 ms out of line; the 2-D loop with two copies is 2.5 % faster inlined).  Same operations on the same values.
		if constexpr (!OUTLINE) { mhdRoeFluxFast<Eqn, SIDE>(F, s, UL, UR); return; }
		typedef typename Eqn::real real;
		constexpr int n = SIDE, t1 = (SIDE + 1) % 3, t2 = (SIDE + 2) % 3;
		VecN<real, 8> a, b;
		a.v[0] = UL[0]; a.v[1] = UL[1 + n]; a.v[2] = UL[1 + t1]; a.v[3] = UL[1 + t2]; a.v[4] = UL[4]; a.v[5] = UL[5 + n]; a.v[6] = UL[5 + t1]; a.v[7] = UL[5 + t2];
		b.v[0] = UR[0]; b.v[1] = UR[1 + n]; b.v[2] = UR[1 + t1]; b.v[3] = UR[1 + t2]; b.v[4] = UR[4]; b.v[5] = UR[5 + n]; b.v[6] = UR[5 + t1]; b.v[7] = UR[5 + t2];
		VecN<real, 8> const r = mhdRoeFluxFastOutOfLine<Eqn>(s, a, b);
		F[0] = r.v[0]; F[1 + n] = r.v[1]; F[1 + t1] = r.v[2]; F[1 + t2] = r.v[3]; F[4] = r.v[4]; F[5 + n] = r.v[5]; F[5 + t1] = r.v[6]; F[5 + t2] = r.v[7];
#else
		mhdRoeFluxFast<Eqn, SIDE>(F, s, UL, UR);
#endif
	}
	else roeFlux<Eqn, SIDE>(F, s, UL, UR);
}

// the literal constrainU + calcDTCell, out of line: cells on the density / pressure floor
template<class Eqn>
HB_NOINLINE void mhdFinishCellOutOfLine(typename Eqn::Params s, typename Eqn::real (&U)[Eqn::nI], typename Eqn::real const (&dx)[3], int dim,
	bool wantDt, typename Eqn::real& dtCell)
{
	typedef typename Eqn::real real;
	Eqn::constrainU(s, U);
	if (wantDt) dtCell = rmin<real>(dtCell, Eqn::calcDTCell(s, U, dx, dim));
}

// constrainU (solverbase.lua:2116-2127 -> euler.cl:698-717) and, when wanted, the cell's CFL rate max_s(lambda_s / dx_s)
// (calcDT.cl:38-73 computes min_s dx_s / max(lambda_s, 1e-9); the kernel inverts the reduced maximum once).
// Production form for Euler: one reciprocal; momentum is left as it is (the reference's m = (m / rho) * rho round trip moves it
// by at most one ulp) and ETotal is rebuilt only when the pressure floor acts.
template<class Eqn>
HB_HD void finishCellAuto(typename Eqn::Params const& s, typename Eqn::real (&U)[Eqn::nI], typename Eqn::real const (&dx)[3],
	typename Eqn::real const (&invdx)[3], int dim, bool wantDt, typename Eqn::real& dtCell, typename Eqn::real& rateCell)
{
	typedef typename Eqn::real real;
	if constexpr (Eqn::FAST && Eqn::eqnId == 0) {
		if (U[0] < s.rhoMin) U[0] = s.rhoMin;
		real const iR = fastRcp(U[0]);
		real const v0 = U[1] * iR, v1 = U[2] * iR, v2 = U[3] * iR;
		real const eK = real(.5) * (U[1] * v0 + U[2] * v1 + U[3] * v2);
		real P = s.gamma_1 * (U[4] - eK);
		if (P < s.PMin) { P = s.PMin; U[4] = eK + P * s.invGamma_1; }
		if (wantDt) {
			real Cs = 0, yCs;
			if (P > s.PMin) fastRsqrt(s.gamma * P * iR, yCs, Cs);
			real r = rmax<real>(rabs(v0) + Cs, real(1e-9)) * invdx[0];
			if (dim > 1) r = rmax<real>(r, rmax<real>(rabs(v1) + Cs, real(1e-9)) * invdx[1]);
			if (dim > 2) r = rmax<real>(r, rmax<real>(rabs(v2) + Cs, real(1e-9)) * invdx[2]);
			rateCell = rmax<real>(rateCell, r);
		}
	} else if constexpr (Eqn::FAST && Eqn::eqnId == 1) {
		// MHD (mhd.cl:915-931, :473-554): one reciprocal; a cell on neither floor is left as it is (the reference's
		// prim -> cons round trip moves it by rounding only); the CFL rate is |v_s| + Cf_s per side from two square roots
		real const rho = U[0], g1 = s.gamma - real(1.), g2 = s.gamma - real(2.);
		real const iR = fastRcp(rho);
		real const v0 = U[1] * iR, v1 = U[2] * iR, v2 = U[3] * iR;
		real const vSq = lenSq3(v0, v1, v2), BSq = lenSq3(U[5], U[6], U[7]);
		real const P = (U[4] - real(.5) * rho * vSq - real(.5) * BSq * s.iMu0) * g1;
		if (!(rho >= real(1e-7) && P >= real(1e-7))) {
			mhdFinishCellOutOfLine<Eqn>(s, U, dx, dim, wantDt, dtCell);
		} else if (wantDt) {
			real const hTotal = real(.5) * vSq + (P * s.gamma * s.iG1 + BSq) * iR;
			real const vv[3] = {v0, v1, v2};
			#pragma unroll
			for (int sd = 0; sd < 3; ++sd) {
				if (sd < dim) {
					real const Bn = U[5 + sd], Bt1 = U[5 + (sd + 1) % 3], Bt2 = U[5 + (sd + 2) % 3];
					real const BPerpSq = Bt1 * Bt1 + Bt2 * Bt2;
					real const CAxSq = Bn * Bn * iR;
					real const hHydro = hTotal - (CAxSq + BPerpSq * iR);
					real const aT = rmax<real>(g1 * (hHydro - real(.5) * vSq) - g2, real(1e-20));
					real const BSPr = (g1 - g2) * BPerpSq * iR;
					real const CAT = CAxSq + BSPr;
					real const dH = real(.5) * (CAT - aT);
					real y0, disc, Cf;
					fastRsqrt0(dH * dH + aT * BSPr, y0, disc);
					fastRsqrt(real(.5) * (CAT + aT) + disc, y0, Cf);
					rateCell = rmax<real>(rateCell, rmax<real>(rabs(vv[sd]) + Cf, real(1e-9)) * invdx[sd]);
				}
			}
		}
	} else {
		Eqn::constrainU(s, U);
		if (wantDt) dtCell = rmin<real>(dtCell, Eqn::calcDTCell(s, U, dx, dim));
	}
}

}   // namespace hb
