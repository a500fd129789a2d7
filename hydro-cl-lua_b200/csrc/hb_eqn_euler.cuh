// hb_eqn_euler.cuh -- compressible Euler device functions for the fused finite-volume stage kernel.
//
// This is the equation plug-in contract of the reference (hydro/eqn/eqn.lua:382-419) for `euler`:
//   primFromCons        hydro/eqn/euler.cl:140-159   (vacuum guard there is dead code: unconditional overwrite :155-157)
//   consFromPrim        hydro/eqn/euler.cl:163-173
//   fluxFromCons        hydro/eqn/euler.cl:273-291
//   eigen_forInterface  hydro/eqn/euler.cl:347-430   (Roe average; literal rhoEpsilon = 1e-5 vacuum branches)
//   eigen_leftTransform hydro/eqn/euler.cl:434-489
//   eigen_rightTransform hydro/eqn/euler.cl:493-542
//   wave speeds         hydro/eqn/euler.lua:309-325  (v_n - Cs, v_n x3, v_n + Cs)
//   calcDTCell          hydro/eqn/cl/calcDT.cl:38-73 + hydro/eqn/euler.lua:346-373
//   constrainU          hydro/eqn/euler.cl:698-717
// State: integrated variables only, U[0..4] = rho, m.x, m.y, m.z, ETotal (ePot, the 6th cons_t field, is
// never read by the flux and is carried untouched in HBM).  The interface normal is a template parameter
// (Cartesian: normal_t = {side}), so the normal_l/u selectors (hydro/coord/coord.lua:2379-2386) resolve at
// compile time and the `x * 0` terms they would produce are simply not emitted (adding an exact zero is a
// no-op in IEEE arithmetic, so results are unchanged).
#pragma once
#include "hb_math.cuh"

namespace hb {

template<class real_, bool FAST_ = false> struct Euler {
	typedef real_ real;
	static constexpr bool FAST = FAST_;                // production arithmetic (hb_roe_fast.cuh) in the marching kernel
	static constexpr int eqnId = 0;
	static constexpr int nS = 6, nI = 5, nW = 5;
	static constexpr bool roeUseFluxFromCons = true;   // hydro/eqn/eqn.lua:46
	struct Params { real gamma, rhoMin, PMin; real gamma_1, invGamma_1, rhoFloor, quarterGamma_1; };   // the last four: production forms only
	static HB_HD Params makeParams(const double* p) {
		return Params{real(p[0]), real(p[1]), real(p[2]), real(p[0] - 1.), real(1. / (p[0] - 1.)), real(p[1] > 1e-5 ? p[1] : 1e-5), real(.25) * real(p[0] - 1.)};
	}

	struct Prim { real rho, v[3], P; };
	struct Eig { real rho, v[3], hTotal, Cs, vSq; };   // vL == v for the identity metric

	// euler.cl:60-66, :76-83
	static HB_HD real calc_P(Params const& s, real const (&U)[nI]) {
		if (U[0] < s.rhoMin) return real(0.);
		real const EKin = real(.5) * lenSq3(U[1], U[2], U[3]) / U[0];
		return (s.gamma - real(1.)) * (U[4] - EKin);
	}
	static HB_HD void primFromCons(Prim& W, Params const& s, real const (&U)[nI]) {
		W.rho = U[0];
		real const invRho = real(1.) / U[0];
		W.v[0] = U[1] * invRho; W.v[1] = U[2] * invRho; W.v[2] = U[3] * invRho;
		W.P = calc_P(s, U);
	}
	static HB_HD void consFromPrim(real (&U)[nI], Params const& s, Prim const& W) {
		U[0] = W.rho;
		U[1] = W.v[0] * W.rho; U[2] = W.v[1] * W.rho; U[3] = W.v[2] * W.rho;
		U[4] = (W.rho * (real(.5) * lenSq3(W.v[0], W.v[1], W.v[2]))) + (W.P / (s.gamma - real(1.)));
	}
	// prim_t as an array in the reference's field order (rho, v.xyz, P): plm.cl casts prim_t and cons_t into each other
	static constexpr bool hasEigenForCell = true;
	static HB_HD void primArray(real (&w)[nI], Params const& s, real const (&U)[nI]) {
		Prim W; primFromCons(W, s, U);
		w[0] = W.rho; w[1] = W.v[0]; w[2] = W.v[1]; w[3] = W.v[2]; w[4] = W.P;
	}
	static HB_HD void consFromPrimArray(real (&U)[nI], Params const& s, real const (&w)[nI]) {
		Prim W; W.rho = w[0]; W.v[0] = w[1]; W.v[1] = w[2]; W.v[2] = w[3]; W.P = w[4];
		consFromPrim(U, s, W);
	}
	// euler.cl:179-195 apply_dU_dW and :201-222 apply_dW_dU on the (rho, v, P) / (rho, m, ETotal) arrays ('plm eig prim' variants only;
	// cartesian: coord_lower is the identity)
	static HB_HD void apply_dU_dW(real (&r)[nI], Params const& s, real const (&WA)[nI], real const (&W)[nI]) {
		r[0] = W[0];
		for (int q = 0; q < 3; ++q) r[1 + q] = WA[1 + q] * W[0] + W[1 + q] * WA[0];
		r[4] = W[0] * real(.5) * dot3(WA[1], WA[2], WA[3], WA[1], WA[2], WA[3]) + WA[0] * dot3(W[1], W[2], W[3], WA[1], WA[2], WA[3])
			+ W[4] / (s.gamma - real(1.));
	}
	static HB_HD void apply_dW_dU(real (&r)[nI], Params const& s, real const (&WA)[nI], real const (&U)[nI]) {
		r[0] = U[0];
		if (U[0] < s.rhoMin) {
			r[1] = r[2] = r[3] = real(0.);
			r[4] = real(0.);
		} else {
			for (int q = 0; q < 3; ++q) r[1 + q] = U[1 + q] * (real(1.) / WA[0]) - WA[1 + q] * (U[0] / WA[0]);
			r[4] = (s.gamma - real(1.)) * (real(.5) * dot3(WA[1], WA[2], WA[3], WA[1], WA[2], WA[3]) * U[0]
				- dot3(U[1], U[2], U[3], WA[1], WA[2], WA[3]) + U[4]);
		}
	}
	// euler.cl:87-94
	static HB_HD real calc_Cs(Params const& s, Prim const& W) {
		if (W.P <= s.PMin) return real(0.);
		if (W.rho < s.rhoMin) return inf_of<real>::v();
		return rsqrt_ieee(s.gamma * W.P / W.rho);
	}

	template<int SIDE> static HB_HD void fluxFromCons(real (&F)[nI], Params const& s, real const (&U)[nI]) {
		Prim W; primFromCons(W, s, U);
		real const v_n = W.v[SIDE];
		F[0] = U[0] * v_n;
		F[1] = U[1] * v_n; F[2] = U[2] * v_n; F[3] = U[3] * v_n;
		F[1 + SIDE] = F[1 + SIDE] + W.P;
		real const HTotal = U[4] + W.P;
		F[4] = HTotal * v_n;
	}

	static HB_HD void eigFromPrim(Eig& r, Params const& s, Prim const& W, real ETotal) {
		r.rho = W.rho;
		r.v[0] = W.v[0]; r.v[1] = W.v[1]; r.v[2] = W.v[2];
		r.vSq = dot3(W.v[0], W.v[1], W.v[2], W.v[0], W.v[1], W.v[2]);
		r.hTotal = (W.P + ETotal) / W.rho;
		r.Cs = calc_Cs(s, W);
	}

	template<int SIDE> static HB_HD void eigen_forInterface(Eig& r, Params const& s, real const (&UL)[nI], real const (&UR)[nI]) {
		real const rhoEpsilon = real(1e-5);
		if (UL[0] < rhoEpsilon && UR[0] < rhoEpsilon) {
			r.rho = 0; r.v[0] = r.v[1] = r.v[2] = 0; r.vSq = 0; r.hTotal = 0; r.Cs = 0;
		} else if (UL[0] < rhoEpsilon) {
			Prim WR; primFromCons(WR, s, UR);
			eigFromPrim(r, s, WR, UR[4]);
		} else if (UR[0] < rhoEpsilon) {
			Prim WL; primFromCons(WL, s, UL);
			eigFromPrim(r, s, WL, UL[4]);
		} else {
			Prim WL; primFromCons(WL, s, UL);
			real const sqrtRhoL = rsqrt_ieee(WL.rho);
			real const hTotalL = (WL.P + UL[4]) / WL.rho;
			Prim WR; primFromCons(WR, s, UR);
			real const sqrtRhoR = rsqrt_ieee(WR.rho);
			real const hTotalR = (WR.P + UR[4]) / WR.rho;
			real const invDenom = real(1.) / (sqrtRhoL + sqrtRhoR);
			r.rho = sqrtRhoL * sqrtRhoR;
			real const wL = sqrtRhoL * invDenom, wR = sqrtRhoR * invDenom;
			r.v[0] = WL.v[0] * wL + WR.v[0] * wR;
			r.v[1] = WL.v[1] * wL + WR.v[1] * wR;
			r.v[2] = WL.v[2] * wL + WR.v[2] * wR;
			real const hTotal = invDenom * (sqrtRhoL * hTotalL + sqrtRhoR * hTotalR);
			real const vSq = dot3(r.v[0], r.v[1], r.v[2], r.v[0], r.v[1], r.v[2]);
			real const eKin = real(.5) * vSq;
			real const h = hTotal - eKin;
			if (h < rhoEpsilon) {
				r.hTotal = eKin; r.Cs = 0;
			} else {
				r.hTotal = hTotal;
				r.Cs = rsqrt_ieee((s.gamma - real(1.)) * h);
			}
			r.vSq = vSq;
		}
	}

	template<int SIDE> static HB_HD void waves(real (&lam)[nW], Params const&, Eig const& e) {
		real const v_n = e.v[SIDE];
		lam[0] = v_n - e.Cs;
		lam[1] = v_n; lam[2] = v_n; lam[3] = v_n;
		lam[4] = v_n + e.Cs;
	}

	template<int SIDE> static HB_HD void leftTransform(real (&r)[nW], Params const& s, Eig const& e, real const (&X)[nI]) {
		if (e.rho < s.rhoMin) {
			for (int j = 0; j < 5; ++j) r[j] = X[j];
			return;
		}
		real vn[3]; rot<SIDE>::fwd(e.v, vn);
		real const denom = real(2.) * e.Cs * e.Cs;
		real const invDenom = real(1.) / denom;
		real const gamma_1 = s.gamma - real(1.);
		// c[q] = -gamma_1 * v_q -/+ Cs * l1_q
		real cm[3], cp[3];
		for (int q = 0; q < 3; ++q) { cm[q] = -gamma_1 * e.v[q]; cp[q] = cm[q]; }
		cm[SIDE] = cm[SIDE] - e.Cs;
		cp[SIDE] = cp[SIDE] + e.Cs;
		real const hk = real(.5) * gamma_1 * e.vSq;
		real const cv = e.Cs * vn[0];
		r[0] = (X[0] * (hk + cv) + X[1] * cm[0] + X[2] * cm[1] + X[3] * cm[2] + X[4] * gamma_1) * invDenom;
		r[1] = (X[0] * (denom - gamma_1 * e.vSq)
				+ X[1] * real(2.) * gamma_1 * e.v[0]
				+ X[2] * real(2.) * gamma_1 * e.v[1]
				+ X[3] * real(2.) * gamma_1 * e.v[2]
				+ X[4] * real(-2.) * gamma_1) * invDenom;
		r[2] = X[0] * -vn[1] + X[1 + (SIDE + 1) % 3];
		r[3] = X[0] * -vn[2] + X[1 + (SIDE + 2) % 3];
		r[4] = (X[0] * (hk - cv) + X[1] * cp[0] + X[2] * cp[1] + X[3] * cp[2] + X[4] * gamma_1) * invDenom;
	}

	template<int SIDE> static HB_HD void rightTransform(real (&r)[nI], Params const& s, Eig const& e, real const (&X)[nW]) {
		if (e.rho < s.rhoMin) {
			for (int j = 0; j < 5; ++j) r[j] = X[j];
			return;
		}
		real vn[3]; rot<SIDE>::fwd(e.v, vn);
		r[0] = X[0] + X[1] + X[4];
		constexpr int q0 = SIDE, q1 = (SIDE + 1) % 3, q2 = (SIDE + 2) % 3;
		r[1 + q0] = X[0] * (e.v[q0] - e.Cs) + X[1] * e.v[q0] + X[4] * (e.v[q0] + e.Cs);
		r[1 + q1] = X[0] * e.v[q1] + X[1] * e.v[q1] + X[2] + X[4] * e.v[q1];
		r[1 + q2] = X[0] * e.v[q2] + X[1] * e.v[q2] + X[3] + X[4] * e.v[q2];
		real const cv = e.Cs * vn[0];
		r[4] = X[0] * (e.hTotal - cv)
			+ X[1] * real(.5) * e.vSq
			+ X[2] * vn[1]
			+ X[3] * vn[2]
			+ X[4] * (e.hTotal + cv);
	}

	static HB_HD void constrainU(Params const& s, real (&U)[nI]) {
		if (U[0] < s.rhoMin) U[0] = s.rhoMin;
		Prim W; primFromCons(W, s, U);
		if (W.P < s.PMin) W.P = s.PMin;
		consFromPrim(U, s, W);
	}

	// euler.cl:319-341 eigen_forCell (used by 'plm athena'): the cell's own eigensystem, Cs from hTotal - eKin (no floors)
	static HB_HD void eigen_forCell(Eig& r, Params const& s, real const (&U)[nI]) {
		Prim W; primFromCons(W, s, U);
		real const vSq = dot3(W.v[0], W.v[1], W.v[2], W.v[0], W.v[1], W.v[2]);
		real const eKin = real(.5) * vSq;
		real const hTotal = (W.P + U[4]) / W.rho;            // calc_hTotal, euler.cl:17-30
		real const CsSq = (s.gamma - real(1.)) * (hTotal - eKin);
		r.rho = W.rho; r.v[0] = W.v[0]; r.v[1] = W.v[1]; r.v[2] = W.v[2]; r.vSq = vSq; r.hTotal = hTotal; r.Cs = rsqrt_ieee(CsSq);
	}

	// eqn.lua:1134-1146 consWaveCodeMinMax with euler.lua:329-336: Cs from the cons state (euler.cl:98-112), v_n = 0 below rhoMin
	template<int SIDE> static HB_HD void consWaveMinMax(real& lmin, real& lmax, Params const& s, real const (&U)[nI]) {
		real Cs;
		real const P = calc_P(s, U);
		if (P <= s.PMin) Cs = real(0.);
		else if (U[0] < s.rhoMin) Cs = inf_of<real>::v();
		else Cs = rsqrt_ieee(s.gamma * P / U[0]);
		Cs *= real(1.);
		real const v_n = U[0] < s.rhoMin ? real(0.) : U[1 + SIDE] / U[0];
		lmin = v_n - Cs; lmax = v_n + Cs;
	}

	// euler.cl:98-112 calc_Cs_fromCons; calcDT.cl:38-73
	static HB_HD real calcDTCell(Params const& s, real const (&U)[nI], real const (&dx)[3], int dim) {
		real Cs;
		real const P = calc_P(s, U);
		if (P <= s.PMin) Cs = real(0.);
		else if (U[0] < s.rhoMin) Cs = inf_of<real>::v();
		else Cs = rsqrt_ieee(s.gamma * P / U[0]);
		real dt = inf_of<real>::v();
		#pragma unroll
		for (int side = 0; side < 3; ++side) {   // static indices: U stays in registers
			if (side < dim && dx[side] > 0) {
				real const v_n = U[0] < s.rhoMin ? real(0.) : U[1 + side] / U[0];
				real const lambdaMin = v_n - Cs, lambdaMax = v_n + Cs;
				real absLambdaMax = rmax<real>(rabs(lambdaMin), rabs(lambdaMax));
				absLambdaMax = rmax<real>(real(1e-9), absLambdaMax);
				dt = rmin<real>(dt, dx[side] / absLambdaMax);
			}
		}
		return dt;
	}

	// mirror boundary: negate m.side (hydro/solver/gridsolver.lua:662-671,741; eqn.lua:366-370)
	static HB_HD bool mirrorFlips(int var, int side) { return var == 1 + side; }
};

}   // namespace hb
