// hb_roe.cuh -- per-interface Roe flux and per-cell PLM slope, equation-generic.
//
//   roeFlux      hydro/flux/roe.cl:17-163  (calcFluxForInterface, Roe): eig = eigen_forInterface(UL,UR);
//                dUe = L.(UR-UL); per wave j: Fe_j = [lambda_j * (L.Uavg)_j | 0] - .5 lambda_j dUe_j (sgn + phi(r_j)(lambda_j dt/dx - sgn));
//                F = R.Fe [+ .5 (F(UL) + F(UR)) when eqn.roeUseFluxFromCons]
//   plmHalfSlope hydro/solver/plm.cl:32-91 ('plm cons'): limited slope sigma of one integrated variable;
//                the cell's face states are U -/+ .5 sigma (this returns .5 * sigma)
//
// Both are pure functions of register operands so they can be unit-tested on the host (tests/host_check.cpp).
#pragma once
#include "hb_math.cuh"

namespace hb {

// 'plm cons' (plm.cl:56-76).  The exact `== 0` tests are load-bearing.
template<class real> HB_HD real plmHalfSlope(int slopeLimiter, real UL, real U, real UR) {
	real const dUR = UR - U;
	real const dUL = U - UL;
	real const dUC = real(.5) * (dUR - dUL);
	real sigma;
	if (dUC >= 0) {
		real const r = dUR == 0 ? real(0) : (dUL / dUR);
		sigma = limiter<real>(slopeLimiter, r) * dUR;
	} else {
		real const r = dUL == 0 ? real(0) : (dUR / dUL);
		sigma = limiter<real>(slopeLimiter, r) * dUL;
	}
	return real(.5) * sigma;
}

// Roe flux without flux limiter (PLM path, or fluxLimiter == 'donor cell'): roe.cl with useFluxLimiter == false.
template<class Eqn, int SIDE>
HB_HD void roeFlux(typename Eqn::real (&F)[Eqn::nI], typename Eqn::Params const& s,
	typename Eqn::real const (&UL)[Eqn::nI], typename Eqn::real const (&UR)[Eqn::nI])
{
	typedef typename Eqn::real real;
	constexpr int nI = Eqn::nI, nW = Eqn::nW;
	typename Eqn::Eig eig;
	Eqn::template eigen_forInterface<SIDE>(eig, s, UL, UR);
	real lam[nW];
	Eqn::template waves<SIDE>(lam, s, eig);
	real fluxEig[nW];
	if (!Eqn::roeUseFluxFromCons) {
		real UAvg[nI];
		for (int j = 0; j < nI; ++j) UAvg[j] = real(.5) * (UL[j] + UR[j]);
		Eqn::template leftTransform<SIDE>(fluxEig, s, eig, UAvg);
	}
	real dU[nI];
	for (int j = 0; j < nI; ++j) dU[j] = UR[j] - UL[j];
	real dUe[nW];
	Eqn::template leftTransform<SIDE>(dUe, s, eig, dU);
	for (int j = 0; j < nW; ++j) {
		real const lambda = lam[j];
		real base = Eqn::roeUseFluxFromCons ? real(0.) : fluxEig[j] * lambda;
		real const sgn = lambda >= 0 ? real(1) : real(-1);
		fluxEig[j] = base - real(.5) * lambda * dUe[j] * sgn;
	}
	Eqn::template rightTransform<SIDE>(F, s, eig, fluxEig);
	if (Eqn::roeUseFluxFromCons) {
		real FL[nI], FR[nI];
		Eqn::template fluxFromCons<SIDE>(FL, s, UL);
		Eqn::template fluxFromCons<SIDE>(FR, s, UR);
		for (int j = 0; j < nI; ++j) F[j] = F[j] + real(.5) * (FL[j] + FR[j]);
	}
}

// Roe flux with flux limiter phi (no PLM): needs the two neighbouring interfaces' states
// (U2L,UL) and (UR,U2R)  (fvsolver.lua:138-155 -> roe.cl:73-80,113-134).
template<class Eqn, int SIDE>
HB_HD void roeFluxLimited(typename Eqn::real (&F)[Eqn::nI], typename Eqn::Params const& s, int fluxLimiter,
	typename Eqn::real dt_dx,
	typename Eqn::real const (&U2L)[Eqn::nI], typename Eqn::real const (&UL)[Eqn::nI],
	typename Eqn::real const (&UR)[Eqn::nI], typename Eqn::real const (&U2R)[Eqn::nI])
{
	typedef typename Eqn::real real;
	constexpr int nI = Eqn::nI, nW = Eqn::nW;
	typename Eqn::Eig eig;
	Eqn::template eigen_forInterface<SIDE>(eig, s, UL, UR);
	real lam[nW];
	Eqn::template waves<SIDE>(lam, s, eig);
	real fluxEig[nW];
	if (!Eqn::roeUseFluxFromCons) {
		real UAvg[nI];
		for (int j = 0; j < nI; ++j) UAvg[j] = real(.5) * (UL[j] + UR[j]);
		Eqn::template leftTransform<SIDE>(fluxEig, s, eig, UAvg);
	}
	real dU[nI], dUl[nI], dUr[nI];
	for (int j = 0; j < nI; ++j) {
		dU[j] = UR[j] - UL[j];
		dUl[j] = UL[j] - U2L[j];
		dUr[j] = U2R[j] - UR[j];
	}
	real dUe[nW], dUeL[nW], dUeR[nW];
	Eqn::template leftTransform<SIDE>(dUe, s, eig, dU);
	{
		typename Eqn::Eig eigL;
		Eqn::template eigen_forInterface<SIDE>(eigL, s, U2L, UL);
		Eqn::template leftTransform<SIDE>(dUeL, s, eigL, dUl);
	}
	{
		typename Eqn::Eig eigR;
		Eqn::template eigen_forInterface<SIDE>(eigR, s, UR, U2R);
		Eqn::template leftTransform<SIDE>(dUeR, s, eigR, dUr);
	}
	for (int j = 0; j < nW; ++j) {
		real const lambda = lam[j];
		real base = Eqn::roeUseFluxFromCons ? real(0.) : fluxEig[j] * lambda;
		real const sgn = lambda >= 0 ? real(1) : real(-1);
		real rEig;
		if (dUe[j] == 0) rEig = 0;
		else if (lambda >= 0) rEig = dUeL[j] / dUe[j];
		else rEig = dUeR[j] / dUe[j];
		real const phi = limiter<real>(fluxLimiter, rEig);
		fluxEig[j] = base - real(.5) * lambda * dUe[j] * (sgn + phi * (lambda * dt_dx - sgn));
	}
	Eqn::template rightTransform<SIDE>(F, s, eig, fluxEig);
	if (Eqn::roeUseFluxFromCons) {
		real FL[nI], FR[nI];
		Eqn::template fluxFromCons<SIDE>(FL, s, UL);
		Eqn::template fluxFromCons<SIDE>(FR, s, UR);
		for (int j = 0; j < nI; ++j) F[j] = F[j] + real(.5) * (FL[j] + FR[j]);
	}
}

}   // namespace hb
