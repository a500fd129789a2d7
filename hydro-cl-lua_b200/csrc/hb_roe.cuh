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

// 'plm prim' (plm.cl:191-253): slopes of the primitive variables with the one-sided ratio r = dWL / dWR (no sign branch), face states
// converted back to conserved variables
template<class Eqn>
HB_HD void plmPrimFaces(typename Eqn::real (&L)[Eqn::nI], typename Eqn::real (&R)[Eqn::nI], typename Eqn::Params const& s, int slopeLimiter,
	typename Eqn::real const (&UL)[Eqn::nI], typename Eqn::real const (&U)[Eqn::nI], typename Eqn::real const (&UR)[Eqn::nI])
{
	typedef typename Eqn::real real;
	constexpr int nI = Eqn::nI;
	real w[nI], wl[nI], wr[nI], nl[nI], nr[nI];
	Eqn::primArray(w, s, U);
	Eqn::primArray(wl, s, UL);
	Eqn::primArray(wr, s, UR);
	#pragma unroll
	for (int j = 0; j < nI; ++j) {
		real const dWR = wr[j] - w[j];
		real const dWL = w[j] - wl[j];
		real const r = dWR == 0 ? real(0) : (dWL / dWR);
		real const sigma = limiter<real>(slopeLimiter, r) * dWR;
		nl[j] = w[j] - real(.5) * sigma;
		nr[j] = w[j] + real(.5) * sigma;
	}
	Eqn::consFromPrimArray(L, s, nl);
	Eqn::consFromPrimArray(R, s, nr);
}

// 'plm cons with flux' (plm.cl:95-187, MUSCL-Hancock): conserved slopes with the one-sided ratio, then both face states moved by
// .5 dt/dx (F(right face state) - F(left face state)) -- added, as the reference has it
template<class Eqn, int SIDE>
HB_HD void plmConsFluxFaces(typename Eqn::real (&L)[Eqn::nI], typename Eqn::real (&R)[Eqn::nI], typename Eqn::Params const& s, int slopeLimiter,
	typename Eqn::real dt_dx, typename Eqn::real const (&UL)[Eqn::nI], typename Eqn::real const (&U)[Eqn::nI], typename Eqn::real const (&UR)[Eqn::nI])
{
	typedef typename Eqn::real real;
	constexpr int nI = Eqn::nI;
	#pragma unroll
	for (int j = 0; j < nI; ++j) {
		real const dUL = U[j] - UL[j];
		real const dUR = UR[j] - U[j];
		real const r = dUR == 0 ? real(0) : (dUL / dUR);
		real const sigma = limiter<real>(slopeLimiter, r) * dUR;
		L[j] = U[j] - real(.5) * sigma;
		R[j] = U[j] + real(.5) * sigma;
	}
	real FL[nI], FR[nI];
	Eqn::template fluxFromCons<SIDE>(FL, s, L);
	Eqn::template fluxFromCons<SIDE>(FR, s, R);
	#pragma unroll
	for (int j = 0; j < nI; ++j) {
		real const dF = FR[j] - FL[j];
		L[j] += real(.5) * dt_dx * dF;
		R[j] += real(.5) * dt_dx * dF;
	}
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

// 'plm athena' (hydro/solver/plm.cl:782-879): face states of one cell along SIDE from slopes of the PRIMITIVE variables limited in the
// characteristic variables of the cell's own eigensystem (the primitive differences go through eigen_leftTransform as if they were
// conserved ones, as in the reference), Athena's monotonicity clamps, back to conserved variables.
// faceOrder 0: as the reference tree assigns them, result->L = cons(Wrv), result->R = cons(Wlv) (plm.cl:877-878);
// faceOrder 1: L = cons(Wlv), R = cons(Wrv) -- the order that reproduces the errors the reference recorded for this scheme
// (tests/test-order/schemes.lua 'plm-athena' rows; tests/test_oracle_kat.py).  Built for equations with eigen_forCell (euler).
template<class real> HB_HD real clSign(real x) { return x > real(0) ? real(1) : (x < real(0) ? real(-1) : x); }   // OpenCL sign()
template<class Eqn, int SIDE>
HB_HD void plmAthenaFaces(typename Eqn::real (&L)[Eqn::nI], typename Eqn::real (&R)[Eqn::nI], typename Eqn::Params const& s,
	typename Eqn::real const (&UL)[Eqn::nI], typename Eqn::real const (&U)[Eqn::nI], typename Eqn::real const (&UR)[Eqn::nI], int faceOrder)
{
	typedef typename Eqn::real real;
	constexpr int nI = Eqn::nI, nW = Eqn::nW;
	typename Eqn::Eig eig;
	Eqn::eigen_forCell(eig, s, U);
	real W[nI], WL[nI], WR[nI];
	Eqn::primArray(W, s, U); Eqn::primArray(WL, s, UL); Eqn::primArray(WR, s, UR);
	real dWL[nI], dWR[nI], dWC[nI], dWG[nI];
	for (int j = 0; j < nI; ++j) {
		dWL[j] = W[j] - WL[j];
		dWR[j] = WR[j] - W[j];
		dWC[j] = real(.5) * (WR[j] - WL[j]);
		dWG[j] = (dWL[j] * dWR[j]) <= real(0.) ? real(0.) : (real(2.) * dWL[j] * dWR[j] / (dWL[j] + dWR[j]));
	}
	real dal[nW], dar[nW], dac[nW], dag[nW], da[nW];
	Eqn::template leftTransform<SIDE>(dal, s, eig, dWL);
	Eqn::template leftTransform<SIDE>(dar, s, eig, dWR);
	Eqn::template leftTransform<SIDE>(dac, s, eig, dWC);
	Eqn::template leftTransform<SIDE>(dag, s, eig, dWG);
	for (int j = 0; j < nW; ++j) {
		da[j] = 0;
		if (dal[j] * dar[j] > 0) {
			real const lim_slope1 = rmin<real>(rabs(dal[j]), rabs(dar[j]));
			real const lim_slope2 = rmin<real>(rabs(dac[j]), rabs(dag[j]));
			da[j] = clSign<real>(dac[j]) * rmin<real>(real(2.) * lim_slope1, lim_slope2);
		}
	}
	real dWm[nI];
	Eqn::template rightTransform<SIDE>(dWm, s, eig, da);
	real Wlv[nI], Wrv[nI];
	for (int j = 0; j < nI; ++j) {
		Wlv[j] = W[j] - real(.5) * dWm[j];
		Wrv[j] = W[j] + real(.5) * dWm[j];
		real const C = Wrv[j] + Wlv[j];
		Wlv[j] = rmax<real>(rmin<real>(W[j], WL[j]), Wlv[j]);
		Wlv[j] = rmin<real>(rmax<real>(W[j], WL[j]), Wlv[j]);
		Wrv[j] = C - Wlv[j];
		Wrv[j] = rmax<real>(rmin<real>(W[j], WR[j]), Wrv[j]);
		Wrv[j] = rmin<real>(rmax<real>(W[j], WR[j]), Wrv[j]);
		Wlv[j] = C - Wrv[j];
	}
	if (faceOrder == 0) { Eqn::consFromPrimArray(L, s, Wrv); Eqn::consFromPrimArray(R, s, Wlv); }
	else { Eqn::consFromPrimArray(L, s, Wlv); Eqn::consFromPrimArray(R, s, Wrv); }
}

// 'plm eig' (plm.cl:256-427, the `#if 1` body): conserved differences on the cell's left eigenvectors, the slope limiter on the
// characteristic ratio dL / dR, each characteristic slope kept only on the side its wave leaves from, back to conserved variables, then
// the half-step flux difference of 'plm cons with flux'.
template<class Eqn, int SIDE>
HB_HD void plmEigFaces(typename Eqn::real (&L)[Eqn::nI], typename Eqn::real (&R)[Eqn::nI], typename Eqn::Params const& s, int slopeLimiter,
	typename Eqn::real dt_dx, typename Eqn::real const (&UL)[Eqn::nI], typename Eqn::real const (&U)[Eqn::nI], typename Eqn::real const (&UR)[Eqn::nI])
{
	typedef typename Eqn::real real;
	constexpr int nI = Eqn::nI, nW = Eqn::nW;
	real dUL[nI], dUR[nI];
	for (int j = 0; j < nI; ++j) {
		dUL[j] = U[j] - UL[j];
		dUR[j] = UR[j] - U[j];
	}
	typename Eqn::Eig eig;
	Eqn::eigen_forCell(eig, s, U);
	real dULEig[nW], dUREig[nW], lam[nW];
	Eqn::template leftTransform<SIDE>(dULEig, s, eig, dUL);
	Eqn::template leftTransform<SIDE>(dUREig, s, eig, dUR);
	Eqn::template waves<SIDE>(lam, s, eig);
	for (int j = 0; j < nW; ++j) {
		real const rEig = dUREig[j] == 0 ? real(0) : (dULEig[j] / dUREig[j]);
		real const sigma = limiter<real>(slopeLimiter, rEig) * dUREig[j];
		dULEig[j] = sigma;
		dUREig[j] = sigma;
		if (lam[j] >= 0) dUREig[j] = 0;
		if (lam[j] <= 0) dULEig[j] = 0;
	}
	real sL[nI], sR[nI];
	Eqn::template rightTransform<SIDE>(sL, s, eig, dULEig);
	Eqn::template rightTransform<SIDE>(sR, s, eig, dUREig);
	for (int j = 0; j < nI; ++j) {
		L[j] = U[j] - real(.5) * sL[j];
		R[j] = U[j] + real(.5) * sR[j];
	}
	real FL[nI], FR[nI];
	Eqn::template fluxFromCons<SIDE>(FL, s, L);
	Eqn::template fluxFromCons<SIDE>(FR, s, R);
	for (int j = 0; j < nI; ++j) {
		real const dF = FR[j] - FL[j];
		L[j] += real(.5) * dt_dx * dF;
		R[j] += real(.5) * dt_dx * dF;
	}
}

// 'plm eig prim' / 'plm eig prim ref' (plm.cl:536-778): primitive differences through dU/dW and the cell's left eigenvectors, the
// symmetric limiter sign(dC) 2 min(|dL|, |dR|, |dC|) on the characteristic differences (as written: the 2 multiplies the central
// difference as well), characteristic tracing over dt -- against the reference state of the fastest wave for REF --, back through the
// right eigenvectors and dW/dU.  faceOrder 0: as the tree assigns them (result->L = the state extrapolated towards +SIDE); 1: exchanged.
template<class Eqn, int SIDE>
HB_HD void plmEigPrimFaces(typename Eqn::real (&L)[Eqn::nI], typename Eqn::real (&R)[Eqn::nI], typename Eqn::Params const& s, bool ref, int faceOrder,
	typename Eqn::real dt_dx, typename Eqn::real const (&UL)[Eqn::nI], typename Eqn::real const (&U)[Eqn::nI], typename Eqn::real const (&UR)[Eqn::nI])
{
	typedef typename Eqn::real real;
	constexpr int nI = Eqn::nI, nW = Eqn::nW;
	real W[nI], WL[nI], WR[nI];
	Eqn::primArray(W, s, U); Eqn::primArray(WL, s, UL); Eqn::primArray(WR, s, UR);
	real dWL[nI], dWR[nI], dWC[nI];
	for (int j = 0; j < nI; ++j) {
		dWL[j] = W[j] - WL[j];
		dWR[j] = WR[j] - W[j];
		dWC[j] = real(.5) * (WR[j] - WL[j]);
	}
	typename Eqn::Eig eig;
	Eqn::eigen_forCell(eig, s, U);
	real tmp[nI], dWLEig[nW], dWREig[nW], dWCEig[nW], dWMEig[nW], lam[nW];
	Eqn::apply_dU_dW(tmp, s, W, dWL); Eqn::template leftTransform<SIDE>(dWLEig, s, eig, tmp);
	Eqn::apply_dU_dW(tmp, s, W, dWR); Eqn::template leftTransform<SIDE>(dWREig, s, eig, tmp);
	Eqn::apply_dU_dW(tmp, s, W, dWC); Eqn::template leftTransform<SIDE>(dWCEig, s, eig, tmp);
	for (int j = 0; j < nW; ++j) {
		dWMEig[j] = dWLEig[j] * dWREig[j] < real(0.) ? real(0.) : (
			(dWCEig[j] >= real(0.) ? real(1.) : real(-1.)) * real(2.) * rmin<real>(rmin<real>(rabs(dWLEig[j]), rabs(dWREig[j])), rabs(dWCEig[j])));
	}
	Eqn::template waves<SIDE>(lam, s, eig);
	real aL[nW], aR[nW], sL[nI], sR[nI], W2L[nI], W2R[nI];
	if (!ref) {
		for (int j = 0; j < nW; ++j) {
			aL[j] = lam[j] < 0 ? real(0) : dWMEig[j] * real(.5) * (real(1.) - lam[j] * dt_dx);
			aR[j] = lam[j] > 0 ? real(0) : dWMEig[j] * real(.5) * (real(1.) + lam[j] * dt_dx);
		}
		Eqn::template rightTransform<SIDE>(tmp, s, eig, aL); Eqn::apply_dW_dU(sL, s, W, tmp);
		Eqn::template rightTransform<SIDE>(tmp, s, eig, aR); Eqn::apply_dW_dU(sR, s, W, tmp);
		for (int j = 0; j < nI; ++j) {
			W2L[j] = W[j] + sL[j];
			W2R[j] = W[j] - sR[j];
		}
	} else {
		real waveMin = rmin<real>(real(0.), lam[0]);          // eigenWaveCodeMinMax: the first and the last wave of the cell's eigensystem
		real waveMax = rmax<real>(real(0.), lam[nW - 1]);
		real dWM[nI], WLRef[nI], WRRef[nI];
		Eqn::template rightTransform<SIDE>(tmp, s, eig, dWMEig); Eqn::apply_dW_dU(dWM, s, W, tmp);
		for (int j = 0; j < nI; ++j) {
			WLRef[j] = W[j] + real(.5) * (real(1.) - dt_dx * waveMax) * dWM[j];
			WRRef[j] = W[j] - real(.5) * (real(1.) + dt_dx * waveMin) * dWM[j];
		}
		for (int j = 0; j < nW; ++j) {
			aL[j] = lam[j] < 0 ? real(0) : (dWMEig[j] * dt_dx * (waveMax - lam[j]));
			aR[j] = lam[j] > 0 ? real(0) : (dWMEig[j] * dt_dx * (waveMin - lam[j]));
		}
		Eqn::template rightTransform<SIDE>(tmp, s, eig, aL); Eqn::apply_dW_dU(sL, s, W, tmp);
		Eqn::template rightTransform<SIDE>(tmp, s, eig, aR); Eqn::apply_dW_dU(sR, s, W, tmp);
		for (int j = 0; j < nI; ++j) {
			W2R[j] = WRRef[j] + real(.5) * sR[j];
			W2L[j] = WLRef[j] + real(.5) * sL[j];
		}
	}
	if (faceOrder == 0) { Eqn::consFromPrimArray(L, s, W2L); Eqn::consFromPrimArray(R, s, W2R); }
	else { Eqn::consFromPrimArray(L, s, W2R); Eqn::consFromPrimArray(R, s, W2L); }
}

// HLL flux, hydro/flux/hll.cl:5-74 with hllCalcWaveMethod = 'Davis direct bounded' (hydro/flux/hll.lua:10):
//   sL = min(lambdaMin(UL), lambdaMin(interface)), sR = max(lambdaMax(UR), lambdaMax(interface)); interface speeds from the Roe-averaged
//   eigensystem (eqn.lua:1108-1120), cell speeds from the cons state (eqn.lua:1134-1146).
template<class Eqn, int SIDE>
HB_HD void hllFlux(typename Eqn::real (&F)[Eqn::nI], typename Eqn::Params const& s,
	typename Eqn::real const (&UL)[Eqn::nI], typename Eqn::real const (&UR)[Eqn::nI])
{
	typedef typename Eqn::real real;
	constexpr int nI = Eqn::nI, nW = Eqn::nW;
	typename Eqn::Eig eigInt;
	Eqn::template eigen_forInterface<SIDE>(eigInt, s, UL, UR);
	real lam[nW];
	Eqn::template waves<SIDE>(lam, s, eigInt);
	real const lambdaIntMin = lam[0], lambdaIntMax = lam[nW - 1];
	real lambdaLMin, lambdaRMax, unused;
	Eqn::template consWaveMinMax<SIDE>(lambdaLMin, unused, s, UL);
	Eqn::template consWaveMinMax<SIDE>(unused, lambdaRMax, s, UR);
	real const sL = rmin<real>(lambdaLMin, lambdaIntMin);
	real const sR = rmax<real>(lambdaRMax, lambdaIntMax);
	for (int j = 0; j < nI; ++j) F[j] = 0;
	if (0 <= sL) {
		Eqn::template fluxFromCons<SIDE>(F, s, UL);
	} else if (sR <= 0) {
		Eqn::template fluxFromCons<SIDE>(F, s, UR);
	} else if (sL <= 0 && 0 <= sR) {
		real FL[nI], FR[nI];
		Eqn::template fluxFromCons<SIDE>(FL, s, UL);
		Eqn::template fluxFromCons<SIDE>(FR, s, UR);
		for (int j = 0; j < nI; ++j) F[j] = (sR * FL[j] - sL * FR[j] + sL * sR * (UR[j] - UL[j])) / (sR - sL);
	}
}

// Rusanov flux, hydro/flux/rusanov.cl:4-33.  The reference's loop over {L, R} assigns lambdaMax in each pass, so the wave speed of the
// RIGHT state is the one that enters the flux; reproduced.
template<class Eqn, int SIDE>
HB_HD void rusanovFlux(typename Eqn::real (&F)[Eqn::nI], typename Eqn::Params const& s,
	typename Eqn::real const (&UL)[Eqn::nI], typename Eqn::real const (&UR)[Eqn::nI])
{
	typedef typename Eqn::real real;
	constexpr int nI = Eqn::nI;
	real lambdaMax;
	{
		real lmin, lmax;
		Eqn::template consWaveMinMax<SIDE>(lmin, lmax, s, UL);
		lambdaMax = rmax<real>(rabs(lmin), rabs(lmax));
	}
	{
		real lmin, lmax;
		Eqn::template consWaveMinMax<SIDE>(lmin, lmax, s, UR);
		lambdaMax = rmax<real>(rabs(lmin), rabs(lmax));
	}
	real FL[nI], FR[nI];
	Eqn::template fluxFromCons<SIDE>(FL, s, UL);
	Eqn::template fluxFromCons<SIDE>(FR, s, UR);
	for (int j = 0; j < nI; ++j) F[j] = real(.5) * (FL[j] + FR[j] - lambdaMax * (UR[j] - UL[j]));
}

// HLLC flux of the Euler equations, hydro/flux/euler-hllc.cl:14-243: 'Davis direct bounded' wave speeds as in hllFlux, contact speed sStar,
// hllcMethod 0 / 1 / 2 (Toro 2012 eqns 38-39 / variation 1 / variation 2; euler-hllc.lua:17 default 2).  State order rho, m.xyz, ETotal.
template<class Eqn, int SIDE>
HB_HD void hllcFlux(typename Eqn::real (&F)[Eqn::nI], typename Eqn::Params const& s, int method,
	typename Eqn::real const (&UL)[Eqn::nI], typename Eqn::real const (&UR)[Eqn::nI])
{
	typedef typename Eqn::real real;
	constexpr int nI = Eqn::nI, nW = Eqn::nW;
	if constexpr (Eqn::eqnId == 0) {
		typename Eqn::Prim WL, WR;
		Eqn::primFromCons(WL, s, UL);
		Eqn::primFromCons(WR, s, UR);
		typename Eqn::Eig eigInt;
		Eqn::template eigen_forInterface<SIDE>(eigInt, s, UL, UR);
		real lam[nW];
		Eqn::template waves<SIDE>(lam, s, eigInt);
		real lambdaLMin, lambdaRMax, unused;
		Eqn::template consWaveMinMax<SIDE>(lambdaLMin, unused, s, UL);
		Eqn::template consWaveMinMax<SIDE>(unused, lambdaRMax, s, UR);
		real const sL = rmin<real>(lambdaLMin, lam[0]);
		real const sR = rmax<real>(lambdaRMax, lam[nW - 1]);
		real vnL[3], vnR[3];
		rot<SIDE>::fwd(WL.v, vnL); rot<SIDE>::fwd(WR.v, vnR);
		real const sStar = (WR.rho * vnR[0] * (sR - vnR[0]) - WL.rho * vnL[0] * (sL - vnL[0]) + WL.P - WR.P)
			/ (WR.rho * (sR - vnR[0]) - WL.rho * (sL - vnL[0]));
		for (int j = 0; j < nI; ++j) F[j] = 0;
		if (0 <= sL) {
			Eqn::template fluxFromCons<SIDE>(F, s, UL);
		} else if ((sL <= real(0.) && real(0.) <= sStar) || (sStar <= real(0.) && real(0.) <= sR)) {
			// the two star regions are the same formulas with (L, sL) or (R, sR)
			bool const left = sL <= real(0.) && real(0.) <= sStar;
			real const (&U)[nI] = left ? UL : UR;
			typename Eqn::Prim const& W = left ? WL : WR;
			real const (&vn)[3] = left ? vnL : vnR;
			real const sK = left ? sL : sR;
			real FK[nI];
			Eqn::template fluxFromCons<SIDE>(FK, s, U);
			if (method == 0) {
				real UStar[nI];
				UStar[0] = U[0] * (sK - vn[0]) / (sK - sStar);
				real const vs[3] = {sStar, vn[1], vn[2]};
				real vStar[3]; rot<SIDE>::inv(vs, vStar);
				UStar[1] = UStar[0] * vStar[0];
				UStar[2] = UStar[0] * vStar[1];
				UStar[3] = UStar[0] * vStar[2];
				UStar[4] = UStar[0] * (U[4] / U[0] + (sStar - vn[0]) * (sStar + W.P / (U[0] * (sK - vn[0]))));
				for (int i = 0; i < nI; ++i) F[i] = FK[i] + sK * (UStar[i] - U[i]);
			} else {
				real const Um_[3] = {U[1], U[2], U[3]}, Fm_[3] = {FK[1], FK[2], FK[3]};
				real Umn[3], Fmn[3]; rot<SIDE>::fwd(Um_, Umn); rot<SIDE>::fwd(Fm_, Fmn);
				real mn[3];
				if (method == 1) {
					F[0] = (sStar * (sK * U[0] - FK[0])) / (sK - sStar);
					mn[0] = (sStar * (sK * Umn[0] - Fmn[0]) + sK * (W.P + W.rho * (sK - vn[0]) * (sStar - vn[0]))) / (sK - sStar);
					mn[1] = (sStar * (sK * Umn[1] - Fmn[1])) / (sK - sStar);
					mn[2] = (sStar * (sK * Umn[2] - Fmn[2])) / (sK - sStar);
					F[4] = (sStar * (sK * U[4] - FK[4]) + sK * (W.P + W.rho * (sK - vn[0]) * (sStar - vn[0])) * sStar) / (sK - sStar);
				} else {
					real const PLR = real(.5) * (WL.P + WR.P + WL.rho * (sL - vnL[0]) * (sStar - vnL[0]) + WR.rho * (sR - vnR[0]) * (sStar - vnR[0]));
					// euler-hllc.cl:196 vs :216: the density flux is written (sL rho - F) sStar on the left, sStar (sR rho - F) on the right
					F[0] = left ? (sK * U[0] - FK[0]) * sStar / (sK - sStar) : sStar * (sK * U[0] - FK[0]) / (sK - sStar);
					mn[0] = left ? ((sK * Umn[0] - Fmn[0]) * sStar + sK * PLR) / (sK - sStar) : (sStar * (sK * Umn[0] - Fmn[0]) + sK * PLR) / (sK - sStar);
					mn[1] = sStar * (sK * Umn[1] - Fmn[1]) / (sK - sStar);
					mn[2] = sStar * (sK * Umn[2] - Fmn[2]) / (sK - sStar);
					F[4] = (sStar * (sK * U[4] - FK[4]) + sK * PLR * sStar) / (sK - sStar);
				}
				real m[3]; rot<SIDE>::inv(mn, m);
				F[1] = m[0]; F[2] = m[1]; F[3] = m[2];
			}
		} else if (sR <= 0) {
			Eqn::template fluxFromCons<SIDE>(F, s, UR);
		} else if (sL <= 0 && 0 <= sR) {
			real FL[nI], FR[nI];
			Eqn::template fluxFromCons<SIDE>(FL, s, UL);
			Eqn::template fluxFromCons<SIDE>(FR, s, UR);
			for (int j = 0; j < nI; ++j) F[j] = (sR * FL[j] - sL * FR[j] + sL * sR * (UR[j] - UL[j])) / (sR - sL);
		}
	} else {
		for (int j = 0; j < nI; ++j) F[j] = 0;       // rejected at hb_fv_create for other equations
	}
}

// the solver's flux plug-in (hydro/flux/*.lua), selected at run time in the tile kernel: 0 roe, 1 hll, 2 rusanov, 3 euler-hllc
template<class Eqn, int SIDE>
HB_HD void interfaceFlux(int flux, int fluxParam, typename Eqn::real (&F)[Eqn::nI], typename Eqn::Params const& s,
	typename Eqn::real const (&UL)[Eqn::nI], typename Eqn::real const (&UR)[Eqn::nI])
{
	if (flux == 1) hllFlux<Eqn, SIDE>(F, s, UL, UR);
	else if (flux == 2) rusanovFlux<Eqn, SIDE>(F, s, UL, UR);
	else if (flux == 3) hllcFlux<Eqn, SIDE>(F, s, fluxParam, UL, UR);
	else roeFlux<Eqn, SIDE>(F, s, UL, UR);
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
