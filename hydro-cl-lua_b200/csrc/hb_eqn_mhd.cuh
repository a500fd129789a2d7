// hb_eqn_mhd.cuh -- ideal MHD (Stone et al. 2008 / Athena 7-wave eigensystem) device functions.
//
// Equation plug-in contract of the reference for `mhd` (file:line into the reference tree):
//   primFromCons / consFromPrim   hydro/eqn/mhd.cl:160-179, :184-201   (floors P, rho at 1e-7 in primFromCons)
//   calcRoeValues                 hydro/eqn/mhd.cl:296-340   (PMag omits mu0; sqrt(rho) weights of B.y, B.z are swapped
//                                                              relative to v -- :335-336 -- and that is reproduced here)
//   eigen_forRoeAvgs              hydro/eqn/mhd.cl:346-436
//   eigen_forInterface            hydro/eqn/mhd.cl:558-571
//   fluxFromCons                  hydro/eqn/mhd.cl:441-467
//   eigen_leftTransform           hydro/eqn/mhd.cl:575-679
//   eigen_rightTransform          hydro/eqn/mhd.cl:683-783
//   wave speeds                   hydro/eqn/mhd.lua:361-373   (v.x -/+ Cf, CAx, Cs ; v.x)
//   calcCellMinMaxEigenvalues     hydro/eqn/mhd.cl:473-554    (the live #else branch, incl. its BStarPerpSq/aTildeSq/hTotal forms)
//   calcDTCell                    hydro/eqn/cl/calcDT.cl:38-73 + hydro/eqn/mhd.lua:377-387
//   constrainU                    hydro/eqn/mhd.cl:915-931
// State: integrated variables U[0..7] = rho, m.xyz, ETotal, B.xyz (psi and ePot are carried untouched in HBM).
// All mu0 uses are solver->mu0 / unit_kg_m_per_C2 (hydro/code/math.cl:270) = Params::mu0.
// Parity contract: the NoDiv / selfgrav ops the reference attaches to this equation are disabled.
#pragma once
#include "hb_math.cuh"

namespace hb {

template<class real_, bool FAST_ = false> struct MHD {
	typedef real_ real;
	static constexpr bool FAST = FAST_;                // production forms: hb_roe_fast.cuh (mhdRoeFluxFast, finishCellAuto)
	static constexpr int eqnId = 1;
	static constexpr int nS = 10, nI = 8, nW = 7;
	static constexpr bool roeUseFluxFromCons = true;   // hydro/eqn/mhd.lua:19
	// l23s: +1 = the reference's left eigenvector entry l23 = .5 betaZ (mhd.cl:621, the parity contract), -1 = Stone et al. 2008's -.5 betaZ
	// (eqn_params[2] < 0; with the reference's sign R L != I whenever both transverse field components are non-zero: tests/test_mhd_alfven.py)
	struct Params { real gamma, mu0; real g2_g1, iMu0, iG1; real l23s; };   // g2_g1, iMu0, iG1: production forms only (gamma_2/gamma_1, 1/mu0, 1/gamma_1)
	static HB_HD Params makeParams(const double* p) {
		return Params{real(p[0]), real(p[1]), real((p[0] - 2.) / (p[0] - 1.)), real(1. / p[1]), real(1. / (p[0] - 1.)), real(p[2] < 0. ? -1. : 1.)};
	}

	struct Prim { real rho, v[3], P, B[3]; };
	struct Eig {
		real rho, v[3], hTotal, B[3], X, Y;
		real hHydro, aTildeSq, Cs, CAx, Cf, BStarPerpLen, betaY, betaZ, betaStarY, betaStarZ, betaStarSq;
		real alphaF, alphaS, sqrtRho, sbx, Qf, Qs, Af, As;
	};

	static HB_HD void primFromCons(Prim& W, Params const& s, real const (&U)[nI]) {
		real const invRho = real(1.) / U[0];
		W.v[0] = U[1] * invRho; W.v[1] = U[2] * invRho; W.v[2] = U[3] * invRho;
		W.B[0] = U[5]; W.B[1] = U[6]; W.B[2] = U[7];
		real const vSq = lenSq3(W.v[0], W.v[1], W.v[2]);
		real const BSq = lenSq3(W.B[0], W.B[1], W.B[2]);
		real const EKin = real(.5) * U[0] * vSq;
		real const EMag = real(.5) * BSq / s.mu0;
		real const EInt = U[4] - EKin - EMag;
		W.P = rmax<real>(EInt * (s.gamma - real(1.)), real(1e-7));
		W.rho = rmax<real>(U[0], real(1e-7));
	}
	static HB_HD void consFromPrim(real (&U)[nI], Params const& s, Prim const& W) {
		U[0] = W.rho;
		U[1] = W.v[0] * W.rho; U[2] = W.v[1] * W.rho; U[3] = W.v[2] * W.rho;
		U[5] = W.B[0]; U[6] = W.B[1]; U[7] = W.B[2];
		real const vSq = lenSq3(W.v[0], W.v[1], W.v[2]);
		real const BSq = lenSq3(W.B[0], W.B[1], W.B[2]);
		real const EKin = real(.5) * W.rho * vSq;
		real const EMag = real(.5) * BSq / s.mu0;
		real const EInt = W.P / (s.gamma - real(1.));
		U[4] = EInt + EKin + EMag;
	}

	template<int SIDE> static HB_HD void fluxFromCons(real (&F)[nI], Params const& s, real const (&U)[nI]) {
		Prim W; primFromCons(W, s, U);
		real const vj = W.v[SIDE];
		real const Bj = W.B[SIDE];
		real const BSq = lenSq3(W.B[0], W.B[1], W.B[2]);
		real const BDotV = dot3(W.B[0], W.B[1], W.B[2], W.v[0], W.v[1], W.v[2]);
		real const PMag = real(.5) * BSq / s.mu0;
		real const PTotal = W.P + PMag;
		real const HTotal = U[4] + PTotal;
		real const Bj_mu0 = Bj / s.mu0;
		F[0] = U[1 + SIDE];
		F[1] = U[1] * vj - U[5] * Bj_mu0;
		F[2] = U[2] * vj - U[6] * Bj_mu0;
		F[3] = U[3] * vj - U[7] * Bj_mu0;
		F[1 + SIDE] = F[1 + SIDE] + PTotal;
		F[5] = U[5] * vj - W.v[0] * Bj;
		F[6] = U[6] * vj - W.v[1] * Bj;
		F[7] = U[7] * vj - W.v[2] * Bj;
		F[4] = HTotal * vj - BDotV * Bj / s.mu0;
	}

	template<int SIDE> static HB_HD void eigen_forInterface(Eig& e, Params const& s, real const (&UL)[nI], real const (&UR)[nI]) {
		// ---- calcRoeValues
		Prim WL; primFromCons(WL, s, UL);
		real const sqrtRhoL = rsqrt_ieee(UL[0]);
		real const PMagL = real(.5) * lenSq3(UL[5], UL[6], UL[7]);
		real const hTotalL = (UL[4] + WL.P + PMagL) / UL[0];
		real vL[3], BL[3]; rot<SIDE>::fwd(WL.v, vL); rot<SIDE>::fwd(WL.B, BL);
		Prim WR; primFromCons(WR, s, UR);
		real const sqrtRhoR = rsqrt_ieee(UR[0]);
		real const PMagR = real(.5) * lenSq3(UR[5], UR[6], UR[7]);
		real const hTotalR = (UR[4] + WR.P + PMagR) / UR[0];
		real vR[3], BR[3]; rot<SIDE>::fwd(WR.v, vR); rot<SIDE>::fwd(WR.B, BR);
		real const dby = BL[1] - BR[1];
		real const dbz = BL[2] - BR[2];
		real const invDenom = real(1.) / (sqrtRhoL + sqrtRhoR);
		real const rho = sqrtRhoL * sqrtRhoR;
		real v[3];
		for (int q = 0; q < 3; ++q) v[q] = (vL[q] * sqrtRhoL + vR[q] * sqrtRhoR) * invDenom;
		real const hTotal = (sqrtRhoL * hTotalL + sqrtRhoR * hTotalR) * invDenom;
		real B[3];
		B[0] = (sqrtRhoL * BL[0] + sqrtRhoR * BR[0]) * invDenom;
		B[1] = (sqrtRhoR * BL[1] + sqrtRhoL * BR[1]) * invDenom;
		B[2] = (sqrtRhoR * BL[2] + sqrtRhoL * BR[2]) * invDenom;
		real const X = real(.5) * (dby * dby + dbz * dbz) * invDenom * invDenom;
		real const Y = real(.5) * (UL[0] + UR[0]) / rho;
		eigen_forRoeAvgs(e, s, rho, v, hTotal, B, X, Y);
	}

	// eigen_forRoeAvgs, hydro/eqn/mhd.cl:346-436: v, B are taken as x-aligned with the interface normal
	static HB_HD void eigen_forRoeAvgs(Eig& e, Params const& s, real rho, real const (&v)[3], real hTotal, real const (&B)[3], real X, real Y) {
		real const gamma_1 = s.gamma - real(1.);
		real const gamma_2 = s.gamma - real(2.);
		real const _1_rho = real(1.) / rho;
		real const vSq = lenSq3(v[0], v[1], v[2]);
		real const BPerpSq = B[1] * B[1] + B[2] * B[2];
		real const BStarPerpSq = (gamma_1 - gamma_2 * Y) * BPerpSq;
		real const CAxSq = B[0] * B[0] * _1_rho;
		real const CASq = CAxSq + BPerpSq * _1_rho;
		e.hHydro = hTotal - CASq;
		e.aTildeSq = rmax<real>((gamma_1 * (e.hHydro - real(.5) * vSq) - gamma_2 * X), real(1e-20));
		real const BStarPerpSq_rho = BStarPerpSq * _1_rho;
		real const CATildeSq = CAxSq + BStarPerpSq_rho;
		real const CStarSq = real(.5) * (CATildeSq + e.aTildeSq);
		real const CA_a_TildeSqDiff = real(.5) * (CATildeSq - e.aTildeSq);
		real const sqrtDiscr = rsqrt_ieee(CA_a_TildeSqDiff * CA_a_TildeSqDiff + e.aTildeSq * BStarPerpSq_rho);
		e.CAx = rsqrt_ieee(CAxSq);
		real const CfSq = CStarSq + sqrtDiscr;
		e.Cf = rsqrt_ieee(CfSq);
		real const CsSq = e.aTildeSq * CAxSq / CfSq;
		e.Cs = rsqrt_ieee(CsSq);
		real const BPerpLen = rsqrt_ieee(BPerpSq);
		e.BStarPerpLen = rsqrt_ieee(BStarPerpSq);
		if (BPerpLen == 0) { e.betaY = 1; e.betaZ = 0; }
		else { e.betaY = B[1] / BPerpLen; e.betaZ = B[2] / BPerpLen; }
		real const sq = rsqrt_ieee(gamma_1 - gamma_2 * Y);
		e.betaStarY = e.betaY / sq;
		e.betaStarZ = e.betaZ / sq;
		e.betaStarSq = e.betaStarY * e.betaStarY + e.betaStarZ * e.betaStarZ;
		if (CfSq - CsSq == 0) { e.alphaF = 1; e.alphaS = 0; }
		else if (e.aTildeSq - CsSq <= 0) { e.alphaF = 0; e.alphaS = 1; }
		else if (CfSq - e.aTildeSq <= 0) { e.alphaF = 1; e.alphaS = 0; }
		else {
			e.alphaF = rsqrt_ieee((e.aTildeSq - CsSq) / (CfSq - CsSq));
			e.alphaS = rsqrt_ieee((CfSq - e.aTildeSq) / (CfSq - CsSq));
		}
		e.sqrtRho = rsqrt_ieee(rho);
		real const _1_sqrtRho = real(1.) / e.sqrtRho;
		e.sbx = B[0] >= 0 ? real(1) : real(-1);
		real const aTilde = rsqrt_ieee(e.aTildeSq);
		e.Qf = e.Cf * e.alphaF * e.sbx;
		e.Qs = e.Cs * e.alphaS * e.sbx;
		e.Af = aTilde * e.alphaF * _1_sqrtRho;
		e.As = aTilde * e.alphaS * _1_sqrtRho;
		e.rho = rho; e.hTotal = hTotal; e.X = X; e.Y = Y;
		for (int q = 0; q < 3; ++q) { e.v[q] = v[q]; e.B[q] = B[q]; }
	}

	// eigen_forCell, hydro/eqn/mhd.cl:859-881 (used by 'plm athena').  The cell's v and B enter as they are, NOT rotated into the
	// normal's frame as calcRoeValues does -- reproduced.
	static constexpr bool hasEigenForCell = true;
	static HB_HD void eigen_forCell(Eig& e, Params const& s, real const (&U)[nI]) {
		Prim W; primFromCons(W, s, U);
		real const PMag = real(.5) * lenSq3(W.B[0], W.B[1], W.B[2]);
		real const hTotal = (U[4] + W.P + PMag) / W.rho;
		eigen_forRoeAvgs(e, s, W.rho, W.v, hTotal, W.B, real(0), real(1));
	}
	// prim_t as an array in the reference's field order (rho, v.xyz, P, B.xyz): plm.cl casts prim_t and cons_t into each other
	static HB_HD void primArray(real (&w)[nI], Params const& s, real const (&U)[nI]) {
		Prim W; primFromCons(W, s, U);
		w[0] = W.rho; w[1] = W.v[0]; w[2] = W.v[1]; w[3] = W.v[2]; w[4] = W.P; w[5] = W.B[0]; w[6] = W.B[1]; w[7] = W.B[2];
	}
	static HB_HD void consFromPrimArray(real (&U)[nI], Params const& s, real const (&w)[nI]) {
		Prim W; W.rho = w[0]; W.v[0] = w[1]; W.v[1] = w[2]; W.v[2] = w[3]; W.P = w[4]; W.B[0] = w[5]; W.B[1] = w[6]; W.B[2] = w[7];
		consFromPrim(U, s, W);
	}

	// mhd.cl:205-223 apply_dU_dW (as written: the result's B is the CELL's B, `(result)->B = (WA)->B`) and :227-245 apply_dW_dU, on the
	// (rho, v, P, B) / (rho, m, ETotal, B) arrays ('plm eig prim' variants only)
	static HB_HD void apply_dU_dW(real (&r)[nI], Params const& s, real const (&WA)[nI], real const (&W)[nI]) {
		r[0] = W[0];
		for (int q = 0; q < 3; ++q) r[1 + q] = WA[1 + q] * W[0] + W[1 + q] * WA[0];
		for (int q = 0; q < 3; ++q) r[5 + q] = WA[5 + q];
		r[4] = W[0] * real(.5) * dot3(WA[1], WA[2], WA[3], WA[1], WA[2], WA[3]) + WA[0] * dot3(W[1], W[2], W[3], WA[1], WA[2], WA[3])
			+ dot3(W[5], W[6], W[7], WA[5], WA[6], WA[7]) / s.mu0 + W[4] / (s.gamma - real(1.));
	}
	static HB_HD void apply_dW_dU(real (&r)[nI], Params const& s, real const (&WA)[nI], real const (&U)[nI]) {
		r[0] = U[0];
		for (int q = 0; q < 3; ++q) r[1 + q] = U[1 + q] * (real(1.) / WA[0]) - WA[1 + q] * (U[0] / WA[0]);
		for (int q = 0; q < 3; ++q) r[5 + q] = U[5 + q];
		r[4] = (s.gamma - real(1.)) * (real(.5) * U[0] * dot3(WA[1], WA[2], WA[3], WA[1], WA[2], WA[3])
			- dot3(U[1], U[2], U[3], WA[1], WA[2], WA[3]) - dot3(U[5], U[6], U[7], WA[5], WA[6], WA[7]) / s.mu0 + U[4]);
	}

	template<int SIDE> static HB_HD void waves(real (&lam)[nW], Params const&, Eig const& e) {
		lam[0] = e.v[0] - e.Cf;
		lam[1] = e.v[0] - e.CAx;
		lam[2] = e.v[0] - e.Cs;
		lam[3] = e.v[0];
		lam[4] = e.v[0] + e.Cs;
		lam[5] = e.v[0] + e.CAx;
		lam[6] = e.v[0] + e.Cf;
	}

	template<int SIDE> static HB_HD void leftTransform(real (&r)[nW], Params const& s, Eig const& e, real const (&U)[nI]) {
		real const Um_[3] = {U[1], U[2], U[3]}, UB_[3] = {U[5], U[6], U[7]};
		real Um[3], UB[3]; rot<SIDE>::fwd(Um_, Um); rot<SIDE>::fwd(UB_, UB);
		real const gamma_1 = s.gamma - real(1.);
		real const gamma_2 = s.gamma - real(2.);
		real const rho = e.rho;
		real const vx = e.v[0], vy = e.v[1], vz = e.v[2];
		real const By = e.B[1], Bz = e.B[2];
		real const X = e.X;
		real const Cs = e.Cs, Cf = e.Cf, BStarPerpLen = e.BStarPerpLen;
		real const betaY = e.betaY, betaZ = e.betaZ, betaStarY = e.betaStarY, betaStarZ = e.betaStarZ, betaStarSq = e.betaStarSq;
		real const alphaF = e.alphaF, alphaS = e.alphaS, sqrtRho = e.sqrtRho, sbx = e.sbx;
		real const Qf = e.Qf, Qs = e.Qs, Af = e.Af, As = e.As;
		real const vSq = lenSq3(vx, vy, vz);
		real const norm = real(.5) / e.aTildeSq;
		real const Cff = norm * alphaF * Cf;
		real const Css = norm * alphaS * Cs;
		real const Qf2 = Qf * norm;
		real const Qs2 = Qs * norm;
		real const AHatF = norm * Af * rho;
		real const AHatS = norm * As * rho;
		real const afpb = norm * Af * BStarPerpLen;
		real const aspb = norm * As * BStarPerpLen;
		real const norm2 = norm * gamma_1;
		real const alphaF2 = alphaF * norm2;
		real const alphaS2 = alphaS * norm2;
		real const QStarY = betaStarY / betaStarSq;
		real const QStarZ = betaStarZ / betaStarSq;
		real const vqstr = (vy * QStarY + vz * QStarZ);
		real const norm3 = norm2 * real(2.);
		real const l16 = AHatS * QStarY - alphaF2 * By;
		real const l17 = AHatS * QStarZ - alphaF2 * Bz;
		real const l21 = real(.5) * (vy * betaZ - vz * betaY);
		real l23 = real(.5) * betaZ;
		if (s.l23s < real(0)) l23 = -l23;
		real const l24 = real(.5) * betaY;
		real const l26 = real(-.5) * sqrtRho * betaZ * sbx;
		real const l27 = real(.5) * sqrtRho * betaY * sbx;
		real const l36 = -AHatF * QStarY - alphaS2 * By;
		real const l37 = -AHatF * QStarZ - alphaS2 * Bz;
		real const Urho = U[0], UE = U[4];
		r[0] =
			  Urho * (alphaF2 * (vSq - e.hHydro) + Cff * (Cf + vx) - Qs2 * vqstr - aspb)
			+ Um[0] * (-alphaF2 * vx - Cff)
			+ Um[1] * (-alphaF2 * vy + Qs2 * QStarY)
			+ Um[2] * (-alphaF2 * vz + Qs2 * QStarZ)
			+ UE * alphaF2
			+ UB[1] * l16
			+ UB[2] * l17;
		r[1] =
			  Urho * l21
			+ Um[1] * l23
			+ Um[2] * l24
			+ UB[1] * l26
			+ UB[2] * l27;
		r[2] =
			  Urho * (alphaS2 * (vSq - e.hHydro) + Css * (Cs + vx) + Qf2 * vqstr + afpb)
			+ Um[0] * (-alphaS2 * vx - Css)
			+ Um[1] * (-alphaS2 * vy - Qf2 * QStarY)
			+ Um[2] * (-alphaS2 * vz - Qf2 * QStarZ)
			+ UE * alphaS2
			+ UB[1] * l36
			+ UB[2] * l37;
		r[3] =
			  Urho * (real(1.) - norm3 * (real(.5) * vSq - gamma_2 * X / gamma_1))
			+ Um[0] * norm3 * vx
			+ Um[1] * norm3 * vy
			+ Um[2] * norm3 * vz
			+ UE * -norm3
			+ UB[1] * norm3 * By
			+ UB[2] * norm3 * Bz;
		r[4] =
			  Urho * (alphaS2 * (vSq - e.hHydro) + Css * (Cs - vx) - Qf2 * vqstr + afpb)
			+ Um[0] * (-alphaS2 * vx + Css)
			+ Um[1] * (-alphaS2 * vy + Qf2 * QStarY)
			+ Um[2] * (-alphaS2 * vz + Qf2 * QStarZ)
			+ UE * alphaS2
			+ UB[1] * l36
			+ UB[2] * l37;
		r[5] =
			  Urho * -l21
			+ Um[1] * -l23
			+ Um[2] * -l24
			+ UB[1] * l26
			+ UB[2] * l27;
		r[6] =
			  Urho * (alphaF2 * (vSq - e.hHydro) + Cff * (Cf - vx) + Qs2 * vqstr - aspb)
			+ Um[0] * (-alphaF2 * vx + Cff)
			+ Um[1] * (-alphaF2 * vy - Qs2 * QStarY)
			+ Um[2] * (-alphaF2 * vz - Qs2 * QStarZ)
			+ UE * alphaF2
			+ UB[1] * l16
			+ UB[2] * l17;
	}

	template<int SIDE> static HB_HD void rightTransform(real (&r)[nI], Params const& s, Eig const& e, real const (&in)[nW]) {
		real const gamma_1 = s.gamma - real(1.);
		real const gamma_2 = s.gamma - real(2.);
		real const vx = e.v[0], vy = e.v[1], vz = e.v[2];
		real const X = e.X;
		real const Cs = e.Cs, Cf = e.Cf, BStarPerpLen = e.BStarPerpLen;
		real const betaY = e.betaY, betaZ = e.betaZ, betaStarY = e.betaStarY, betaStarZ = e.betaStarZ, betaStarSq = e.betaStarSq;
		real const alphaF = e.alphaF, alphaS = e.alphaS, sbx = e.sbx;
		real const Qf = e.Qf, Qs = e.Qs, Af = e.Af, As = e.As;
		real const vSq = lenSq3(vx, vy, vz);
		real const vDotBeta = vy * betaStarY + vz * betaStarZ;
		real const _1_sqrtRho = real(1.) / e.sqrtRho;
		real const Afpbb = Af * BStarPerpLen * betaStarSq;
		real const Aspbb = As * BStarPerpLen * betaStarSq;
		real const lambdaFastMin = vx - Cf;
		real const lambdaSlowMin = vx - Cs;
		real const lambdaSlowMax = vx + Cs;
		real const lambdaFastMax = vx + Cf;
		real const qa3 = alphaF * vy;
		real const qb3 = alphaS * vy;
		real const qc3 = Qs * betaStarY;
		real const qd3 = Qf * betaStarY;
		real const qa4 = alphaF * vz;
		real const qb4 = alphaS * vz;
		real const qc4 = Qs * betaStarZ;
		real const qd4 = Qf * betaStarZ;
		real const r52 = -(vy * betaZ - vz * betaY);
		real const r61 = As * betaStarY;
		real const r62 = -betaZ * sbx * _1_sqrtRho;
		real const r63 = -Af * betaStarY;
		real const r71 = As * betaStarZ;
		real const r72 = betaY * sbx * _1_sqrtRho;
		real const r73 = -Af * betaStarZ;
		r[0] =
			  in[0] * alphaF
			+ in[2] * alphaS
			+ in[3]
			+ in[4] * alphaS
			+ in[6] * alphaF;
		real rm[3];
		rm[0] =
			  in[0] * alphaF * lambdaFastMin
			+ in[2] * alphaS * lambdaSlowMin
			+ in[3] * vx
			+ in[4] * alphaS * lambdaSlowMax
			+ in[6] * alphaF * lambdaFastMax;
		rm[1] =
			  in[0] * (qa3 + qc3)
			+ in[1] * -betaZ
			+ in[2] * (qb3 - qd3)
			+ in[3] * vy
			+ in[4] * (qb3 + qd3)
			+ in[5] * betaZ
			+ in[6] * (qa3 - qc3);
		rm[2] =
			  in[0] * (qa4 + qc4)
			+ in[1] * betaY
			+ in[2] * (qb4 - qd4)
			+ in[3] * vz
			+ in[4] * (qb4 + qd4)
			+ in[5] * -betaY
			+ in[6] * (qa4 - qc4);
		real om[3]; rot<SIDE>::inv(rm, om);
		r[1] = om[0]; r[2] = om[1]; r[3] = om[2];
		r[4] =
			  in[0] * (alphaF * (e.hHydro - vx * Cf) + Qs * vDotBeta + Aspbb)
			+ in[1] * r52
			+ in[2] * (alphaS * (e.hHydro - vx * Cs) - Qf * vDotBeta - Afpbb)
			+ in[3] * (real(.5) * vSq + gamma_2 * X / gamma_1)
			+ in[4] * (alphaS * (e.hHydro + vx * Cs) + Qf * vDotBeta - Afpbb)
			+ in[5] * -r52
			+ in[6] * (alphaF * (e.hHydro + vx * Cf) - Qs * vDotBeta + Aspbb);
		real rB[3];
		rB[0] = 0;
		rB[1] =
			  in[0] * r61
			+ in[1] * r62
			+ in[2] * r63
			+ in[4] * r63
			+ in[5] * r62
			+ in[6] * r61;
		rB[2] =
			  in[0] * r71
			+ in[1] * r72
			+ in[2] * r73
			+ in[4] * r73
			+ in[5] * r72
			+ in[6] * r71;
		real oB[3]; rot<SIDE>::inv(rB, oB);
		r[5] = oB[0]; r[6] = oB[1]; r[7] = oB[2];
	}

	static HB_HD void constrainU(Params const& s, real (&U)[nI]) {
		Prim W; primFromCons(W, s, U);
		W.rho = rmax<real>(W.rho, real(1e-7));
		W.P = rmax<real>(W.P, real(1e-7));
		consFromPrim(U, s, W);
	}

	template<int SIDE> static HB_HD void cellMinMaxEigenvalues(real& lmin, real& lmax, Params const& s, Prim const& W) {
		real const v_n = W.v[SIDE];
		real Bn[3]; rot<SIDE>::fwd(W.B, Bn);
		real const gamma = s.gamma;
		real const gamma_1 = gamma - real(1.);
		real const gamma_2 = gamma - real(2.);
		real const vSq = lenSq3(W.v[0], W.v[1], W.v[2]);
		real const BSq = lenSq3(W.B[0], W.B[1], W.B[2]);
		real const hTotal = real(.5) * vSq + (W.P * gamma / gamma_1 + BSq) / W.rho;
		real const _1_rho = real(1.) / W.rho;
		real const BPerpSq = Bn[1] * Bn[1] + Bn[2] * Bn[2];
		real const BStarPerpSq = (gamma_1 - gamma_2) * BPerpSq;
		real const CAxSq = Bn[0] * Bn[0] * _1_rho;
		real const CASq = CAxSq + BPerpSq * _1_rho;
		real const hHydro = hTotal - CASq;
		real const aTildeSq = rmax<real>((gamma_1 * (hHydro - real(.5) * vSq) - gamma_2), real(1e-20));
		real const BStarPerpSq_rho = BStarPerpSq * _1_rho;
		real const CATildeSq = CAxSq + BStarPerpSq_rho;
		real const CStarSq = real(.5) * (CATildeSq + aTildeSq);
		real const CA_a_TildeSqDiff = real(.5) * (CATildeSq - aTildeSq);
		real const sqrtDiscr = rsqrt_ieee(CA_a_TildeSqDiff * CA_a_TildeSqDiff + aTildeSq * BStarPerpSq_rho);
		real const CfSq = CStarSq + sqrtDiscr;
		real const Cf = rsqrt_ieee(CfSq);
		lmin = v_n - Cf;
		lmax = v_n + Cf;
	}

	// mhd.lua:377-387 consWaveCodeMinMax -> calcCellMinMaxEigenvalues
	template<int SIDE> static HB_HD void consWaveMinMax(real& lmin, real& lmax, Params const& s, real const (&U)[nI]) {
		Prim W; primFromCons(W, s, U);
		cellMinMaxEigenvalues<SIDE>(lmin, lmax, s, W);
	}

	static HB_HD real calcDTCell(Params const& s, real const (&U)[nI], real const (&dx)[3], int dim) {
		real dt = inf_of<real>::v();
		Prim W; primFromCons(W, s, U);
		real lmin, lmax;
		if (dim > 0 && dx[0] > 0) {
			cellMinMaxEigenvalues<0>(lmin, lmax, s, W);
			dt = rmin<real>(dt, dx[0] / rmax<real>(real(1e-9), rmax<real>(rabs(lmin), rabs(lmax))));
		}
		if (dim > 1 && dx[1] > 0) {
			cellMinMaxEigenvalues<1>(lmin, lmax, s, W);
			dt = rmin<real>(dt, dx[1] / rmax<real>(real(1e-9), rmax<real>(rabs(lmin), rabs(lmax))));
		}
		if (dim > 2 && dx[2] > 0) {
			cellMinMaxEigenvalues<2>(lmin, lmax, s, W);
			dt = rmin<real>(dt, dx[2] / rmax<real>(real(1e-9), rmax<real>(rabs(lmin), rabs(lmax))));
		}
		return dt;
	}

	// mirror boundary: negate m.side and B.side (hydro/eqn/eqn.lua:366-370 reflectVars)
	static HB_HD bool mirrorFlips(int var, int side) { return var == 1 + side || var == 5 + side; }
};

}   // namespace hb
