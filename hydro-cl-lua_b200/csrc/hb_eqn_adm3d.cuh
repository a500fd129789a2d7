// hb_eqn_adm3d.cuh -- ADM Bona-Masso 3-D (first-order 3+1 numerical relativity) device functions.
//
// The equation plug-in contract of the reference (hydro/eqn/eqn.lua:382-419) for `adm3d` as BASELINE config C5 runs it:
// noZeroRowsInFlux = true (13 waves per side: a_side, d_side,ij, K_ij), useShift = 'none', roeUseFluxFromCons = false
// (hydro/eqn/adm3d.lua:21,124-143), 'constrain V' = none.
//   eigen_forInterface    hydro/eqn/adm3d.cl:373-397   (arithmetic mean of alpha and gamma_ij)
//   wave speeds           hydro/eqn/adm3d.lua:427-479  (-+ alpha sqrt(f gamma^jj), -+ alpha sqrt(gamma^jj) x5, 0)
//   eigen_leftTransform   hydro/eqn/adm3d.cl:670-707
//   eigen_rightTransform  hydro/eqn/adm3d.cl:712-722,1225-1274
//   calcDTCell            hydro/eqn/cl/calcDT.cl:38-73 + hydro/eqn/adm3d.lua:500-544
//   initDerivs            hydro/eqn/adm3d.cl:196-243   (centred differences exactly as written there: no 1/2 on alpha_,i)
//   addSource             hydro/eqn/adm3d.cl:1402-3072.  The reference's body is a machine-generated polynomial
//                         (:1589-2776); here it is the equivalent tensor form
//       d/dt alpha    += -alpha^2 f K                         d/dt gamma_ij += -2 alpha K_ij
//       d/dt a_k      += -(alpha f + alpha^2 f') a_k K + 2 alpha f K^ij d_kij
//       d/dt d_kij    += -alpha a_k K_ij
//       d/dt K_ij     += alpha (-a_i a_j + conn^k_ij (a_k + d_k - 2 e_k) + d_ikl d_j^kl + 2 d_ki^l (d^k_jl - d_lj^k)
//                               + K K_ij - 2 K_ik K^k_j - 8 pi S_ij + 4 pi gamma_ij (S - rho))
//       d/dt V_k      += V_convCoeff (d_k - e_k - V_k)
//     (conn^k_ij = d_ij^k + d_ji^k - d^k_ij, d_k = d_km^m, e_k = d^m_mk), checked against the reference's own generated lines
//     by tests/test_adm3d.py through oracle/_ref.  a_convCoeff / d_convCoeff (radius-1 terms, default 0) are applied by the kernel.
// State order (37 integrated + 14 carried): alpha, gamma_ll[xx xy xz yy yz zz], a_l[3], d_lll[3][6], K_ll[6], V_l[3] | rho, S_u[3],
// S_ll[6], H, M_u[3].  constrainU only refreshes the diagnostics H, M_u in the reference: they stay zero here.
#pragma once
#include "hb_math.cuh"

namespace hb {

template<class real_, bool FAST_ = false> struct ADM3D {
	typedef real_ real;
	static constexpr bool FAST = FAST_;
	static constexpr int eqnId = 2;
	static constexpr int nS = 51, nI = 37, nW = 13;
	static constexpr bool roeUseFluxFromCons = false;
	enum { iAlpha = 0, iGamma = 1, iA = 7, iD = 10, iK = 28, iV = 34, iRho = 37, iSu = 38, iSll = 41, iH = 47, iMu = 48 };
	struct Params { int f_eqn; real a_conv, d_conv, V_conv; };
	static HB_HD Params makeParams(const double* p) { return Params{int(p[0]), real(p[1]), real(p[2]), real(p[3])}; }

	// what one side's 13-wave system sees of a cell: alpha, gamma_ll, a_side, d_side,ij, K_ij
	struct Side { real alpha, g[6], a, d[6], K[6]; };
	struct Eig { real alpha, alpha_sqrt_f, gU[6], sq[3]; real iAsf, iSq; };   // iAsf = 1 / alpha_sqrt_f, iSq = 1 / sq[SIDE]: production forms only

	static HB_HD int s6(int i, int j) { return i == j ? (i == 0 ? 0 : i == 1 ? 3 : 5) : (i + j == 1 ? 1 : i + j == 2 ? 2 : 4); }

	// gauge function (hydro/init/einstein.lua:54-81; options hydro/eqn/einstein.lua:42-48)
	static HB_HD real fConst(int id) { return id == 2 ? real(1.) : id == 3 ? real(0.) : id == 4 ? real(.49) : id == 5 ? real(.5) : id == 6 ? real(1.5) : real(1.69); }
	static HB_HD real calc_f_alpha(int id, real a) { return id == 0 ? real(2.) : id == 1 ? a + real(1.) / a : fConst(id) * a; }
	static HB_HD real calc_f_alphaSq(int id, real a) { return id == 0 ? real(2.) * a : id == 1 ? a * a + real(1.) : fConst(id) * a * a; }
	static HB_HD real calc_alphaSq_dalpha_f(int id, real a) { return id == 0 ? real(-2.) : id == 1 ? real(-2.) / a : real(0.); }

	// hydro/code/math.cl:518-536,591-594,683-685
	static HB_HD real det6(const real* m) {
		return m[0] * (m[3] * m[5] - m[4] * m[4]) - m[1] * (m[1] * m[5] - m[2] * m[4]) + m[2] * (m[1] * m[4] - m[3] * m[2]);
	}
	static HB_HD void inv6(real* o, const real* m, real det) {
		real const invDet = FAST ? fastRcp(det) : real(1.) / det;
		o[0] = (m[3] * m[5] - m[4] * m[4]) * invDet;
		o[1] = (m[2] * m[4] - m[1] * m[5]) * invDet;
		o[2] = (m[1] * m[4] - m[2] * m[3]) * invDet;
		o[3] = (m[0] * m[5] - m[2] * m[2]) * invDet;
		o[4] = (m[2] * m[1] - m[0] * m[4]) * invDet;
		o[5] = (m[0] * m[3] - m[1] * m[1]) * invDet;
	}
	static HB_HD real dot6(const real* a, const real* b) {
		return a[0] * b[0] + a[3] * b[3] + a[5] * b[5] + real(2.) * (a[1] * b[1] + a[2] * b[2] + a[4] * b[4]);
	}
	// real3s3_swap<SIDE>: exchange the x axis with axis SIDE
	template<int SIDE> static HB_HD void swap6(real* o, const real* m) {
		if (SIDE == 0) { o[0] = m[0]; o[1] = m[1]; o[2] = m[2]; o[3] = m[3]; o[4] = m[4]; o[5] = m[5]; }
		else if (SIDE == 1) { o[0] = m[3]; o[1] = m[1]; o[2] = m[4]; o[3] = m[0]; o[4] = m[2]; o[5] = m[5]; }
		else { o[0] = m[5]; o[1] = m[4]; o[2] = m[2]; o[3] = m[3]; o[4] = m[1]; o[5] = m[0]; }
	}

	// production form: the side's own sqrt(gamma^jj) only (the other two are never read), reciprocals from the same seeds
	template<int SIDE> static HB_HD void eigen_forInterfaceFast(Eig& e, Params const& s, Side const& UL, Side const& UR) {
		e.alpha = real(.5) * (UL.alpha + UR.alpha);
		real avg[6];
		#pragma unroll
		for (int k = 0; k < 6; ++k) avg[k] = (UL.g[k] + UR.g[k]) * real(.5);
		inv6(e.gU, avg, det6(avg));
		fastRsqrt(calc_f_alphaSq(s.f_eqn, e.alpha), e.iAsf, e.alpha_sqrt_f);
		fastRsqrt(e.gU[SIDE == 0 ? 0 : SIDE == 1 ? 3 : 5], e.iSq, e.sq[SIDE]);
	}
	static HB_HD void eigen_forInterface(Eig& e, Params const& s, Side const& UL, Side const& UR) {
		e.alpha = real(.5) * (UL.alpha + UR.alpha);
		real avg[6];
		#pragma unroll
		for (int k = 0; k < 6; ++k) avg[k] = (UL.g[k] + UR.g[k]) * real(.5);
		real const det = det6(avg);
		e.alpha_sqrt_f = rsqrt_ieee(calc_f_alphaSq(s.f_eqn, e.alpha));
		inv6(e.gU, avg, det);
		e.sq[0] = rsqrt_ieee(e.gU[0]);
		e.sq[1] = rsqrt_ieee(e.gU[3]);
		e.sq[2] = rsqrt_ieee(e.gU[5]);
	}
	template<int SIDE> static HB_HD void waves(real (&lam)[nW], Eig const& e) {
		real const sq = e.sq[SIDE];
		real const lambdaLight = sq * e.alpha;
		real const lambdaGauge = sq * e.alpha_sqrt_f;
		lam[0] = -real(0) - lambdaGauge;
		#pragma unroll
		for (int k = 1; k <= 5; ++k) lam[k] = -real(0) - lambdaLight;
		lam[6] = -real(0);
		#pragma unroll
		for (int k = 7; k <= 11; ++k) lam[k] = -real(0) + lambdaLight;
		lam[12] = -real(0) + lambdaGauge;
	}
	// inputs: a_side, d_side,ij and K_ij in the state's own (unswapped) component order
	template<int SIDE> static HB_HD void leftTransform(real (&r)[nW], Eig const& e, real a_j, const real* dIn, const real* KIn) {
		real const _1_sqrt_f = FAST ? e.alpha * e.iAsf : e.alpha / e.alpha_sqrt_f;
		real const _1_f = _1_sqrt_f * _1_sqrt_f;
		real const sqrt_gammaUjj = e.sq[SIDE];
		real const _1_gammaUjj = FAST ? e.iSq * e.iSq : real(1.) / e.gU[SIDE == 0 ? 0 : SIDE == 1 ? 3 : 5];
		real d[6], K[6], gU[6];
		swap6<SIDE>(d, dIn); swap6<SIDE>(K, KIn); swap6<SIDE>(gU, e.gU);
		real const K_dot_eig_gamma = dot6(K, gU);
		real const dj_dot_eig_gamma = dot6(d, gU);
		r[0] = (a_j * -sqrt_gammaUjj * _1_sqrt_f + K_dot_eig_gamma) * real(.5) * _1_gammaUjj;
		#pragma unroll
		for (int i = 1; i <= 5; ++i) r[i] = real(.5) * (-sqrt_gammaUjj * d[i] + K[i]);
		r[6] = (-a_j * _1_f + dj_dot_eig_gamma) * _1_gammaUjj;
		#pragma unroll
		for (int i = 1; i <= 5; ++i) r[6 + i] = real(.5) * (sqrt_gammaUjj * d[i] + K[i]);
		r[12] = (a_j * sqrt_gammaUjj * _1_sqrt_f + K_dot_eig_gamma) * real(.5) * _1_gammaUjj;
	}
	template<int SIDE> static HB_HD void rightTransform(real& aOut, real* dOut, real* KOut, Eig const& e, real const (&in)[nW]) {
		real gU[6];
		swap6<SIDE>(gU, e.gU);
		real const input1_dot_gammaU = in[1] * real(2.) * gU[1] + in[2] * real(2.) * gU[2] + in[3] * gU[3] + in[4] * real(2.) * gU[4] + in[5] * gU[5];
		real const input7_dot_gammaU = in[7] * real(2.) * gU[1] + in[8] * real(2.) * gU[2] + in[9] * gU[3] + in[10] * real(2.) * gU[4] + in[11] * gU[5];
		real const sqrt_f = FAST ? e.alpha_sqrt_f * fastRcp(e.alpha) : e.alpha_sqrt_f / e.alpha;
		real const _1_sqrt_f = FAST ? e.alpha * e.iAsf : real(1.) / sqrt_f;
		real const sqrt_gammaUjj = e.sq[SIDE];
		real const _1_sqrt_gammaUjj = FAST ? e.iSq : real(1.) / sqrt_gammaUjj;
		real const _1_gammaUjj = _1_sqrt_gammaUjj * _1_sqrt_gammaUjj;
		aOut = sqrt_f * sqrt_gammaUjj * (in[12] - in[0]);
		real d[6], K[6];
		d[0] = ((in[12] - in[0]) * _1_sqrt_f + (input1_dot_gammaU - input7_dot_gammaU) * _1_gammaUjj) * _1_sqrt_gammaUjj + in[6];
		#pragma unroll
		for (int i = 1; i <= 5; ++i) d[i] = (in[i + 6] - in[i]) * _1_sqrt_gammaUjj;
		K[0] = in[0] + in[12] - (input1_dot_gammaU + input7_dot_gammaU) * _1_gammaUjj;
		#pragma unroll
		for (int i = 1; i <= 5; ++i) K[i] = in[i] + in[i + 6];
		swap6<SIDE>(dOut, d); swap6<SIDE>(KOut, K);
	}

	// ---- the Roe flux with flux limiter (hydro/flux/roe.cl:17-163, roeUseFluxFromCons == false) of the 13-wave system, in two pieces so
	// that a kernel can share the characteristic differences of an interface with the two interfaces next to it:
	//   interfaceChar  eig = eigen_forInterface(UL, UR); base = L (UL + UR)/2; dUe = L (UR - UL)           (roe.cl:40-72)
	//   limitedFlux    per wave: base_j lambda_j - .5 lambda_j dUe_j (sgn + phi(r_j)(lambda_j dt/dx - sgn)),
	//                  r_j = dUe_j of the upwind neighbour interface / dUe_j; then F = R (...)                 (roe.cl:73-163)
	template<int SIDE> static HB_HD void interfaceEig(Eig& eig, Params const& s, Side const& UL, Side const& UR) {
		if constexpr (FAST) eigen_forInterfaceFast<SIDE>(eig, s, UL, UR); else eigen_forInterface(eig, s, UL, UR);
	}
	template<int SIDE> static HB_HD void charAvg(real (&base)[nW], Eig const& eig, Side const& UL, Side const& UR) {
		real dA[6], KA[6];
		#pragma unroll
		for (int k = 0; k < 6; ++k) { dA[k] = real(.5) * (UL.d[k] + UR.d[k]); KA[k] = real(.5) * (UL.K[k] + UR.K[k]); }
		leftTransform<SIDE>(base, eig, real(.5) * (UL.a + UR.a), dA, KA);
	}
	template<int SIDE> static HB_HD void charDiff(real (&dUe)[nW], Eig const& eig, Side const& UL, Side const& UR) {
		real dd[6], dK[6];
		#pragma unroll
		for (int k = 0; k < 6; ++k) { dd[k] = UR.d[k] - UL.d[k]; dK[k] = UR.K[k] - UL.K[k]; }
		leftTransform<SIDE>(dUe, eig, UR.a - UL.a, dd, dK);
	}
	// fluxEig: in = L (UL + UR)/2, out = the limited characteristic flux
	template<int SIDE> static HB_HD void limitedFlux(real& Fa, real* Fd, real* FK, Eig const& eig, real (&fluxEig)[nW], real const (&dUe)[nW],
		real const (&dUeL)[nW], real const (&dUeR)[nW], int fluxLimiter, bool useLimiter, real dt_dx)
	{
		real lam[nW];
		waves<SIDE>(lam, eig);
		#pragma unroll
		for (int j = 0; j < nW; ++j) {
			real const lambda = lam[j];
			real const base = fluxEig[j] * lambda;
			real const sgn = lambda >= 0 ? real(1) : real(-1);
			if (FAST && useLimiter && (fluxLimiter == 8 || fluxLimiter == 18)) {
				// minmod / superbee without the ratio: phi(r) dUe = sign(dUe) m(|dUe|, |up|) when up dUe > 0, else 0 (hydro/app.lua:622,632)
				real const up = lambda >= 0 ? dUeL[j] : dUeR[j];
				real const a = rabs(dUe[j]), b = rabs(up);
				real const m = fluxLimiter == 8 ? rmin<real>(a, b) : rmax<real>(rmin<real>(a, real(2.) * b), rmin<real>(real(2.) * a, b));
				real const phiD = up * dUe[j] > real(0) ? (dUe[j] > real(0) ? m : -m) : real(0);
				fluxEig[j] = base - real(.5) * lambda * (dUe[j] * sgn + phiD * (lambda * dt_dx - sgn));
			} else if (useLimiter) {
				real rEig;
				if (dUe[j] == 0) rEig = 0;
				else if (lambda >= 0) rEig = dUeL[j] / dUe[j];
				else rEig = dUeR[j] / dUe[j];
				real const phi = limiter<real>(fluxLimiter, rEig);
				fluxEig[j] = base - real(.5) * lambda * dUe[j] * (sgn + phi * (lambda * dt_dx - sgn));
			} else {
				fluxEig[j] = base - real(.5) * lambda * dUe[j] * sgn;
			}
		}
		rightTransform<SIDE>(Fa, Fd, FK, eig, fluxEig);
	}
	// one interface on its own (every characteristic difference computed here): U2L, U2R are the next cells outwards
	template<int SIDE> static HB_HD void roeFluxLimited(real& Fa, real* Fd, real* FK, Params const& s, int fluxLimiter, bool useLimiter, real dt_dx,
		Side const& U2L, Side const& UL, Side const& UR, Side const& U2R)
	{
		Eig eig;
		interfaceEig<SIDE>(eig, s, UL, UR);
		real fluxEig[nW], dUe[nW], dUeL[nW], dUeR[nW];
		charAvg<SIDE>(fluxEig, eig, UL, UR);
		charDiff<SIDE>(dUe, eig, UL, UR);
		if (useLimiter) {
			Eig eL;
			interfaceEig<SIDE>(eL, s, U2L, UL);
			charDiff<SIDE>(dUeL, eL, U2L, UL);
			Eig eR;
			interfaceEig<SIDE>(eR, s, UR, U2R);
			charDiff<SIDE>(dUeR, eR, UR, U2R);
		}
		limitedFlux<SIDE>(Fa, Fd, FK, eig, fluxEig, dUe, dUeL, dUeR, fluxLimiter, useLimiter, dt_dx);
	}

	// U: the cell's 37 integrated variables; rho, Sll: matter terms (zero in the configs)
	static HB_HD void addSource(real (&deriv)[nI], Params const& s, real const (&U)[nI], real rho, const real* Sll) {
		real const alpha = U[iAlpha];
		const real* g = U + iGamma; const real* a = U + iA; const real* K = U + iK; const real* V = U + iV;
		real gU[6];
		inv6(gU, g, det6(g));
		real const Str = dot6(Sll, gU);
		real K_ul[3][3], K_uu[3][3];
		#pragma unroll
		for (int i = 0; i < 3; ++i)
			#pragma unroll
			for (int j = 0; j < 3; ++j) { real t = 0; for (int m = 0; m < 3; ++m) t += gU[s6(i, m)] * K[s6(m, j)]; K_ul[i][j] = t; }
		#pragma unroll
		for (int i = 0; i < 3; ++i)
			#pragma unroll
			for (int j = 0; j < 3; ++j) { real t = 0; for (int m = 0; m < 3; ++m) t += K_ul[i][m] * gU[s6(m, j)]; K_uu[i][j] = t; }
		real const trK = K_ul[0][0] + K_ul[1][1] + K_ul[2][2];
		real d_llu[3][3][3], d_ull[3][3][3], d_luu[3][3][3];
		#pragma unroll
		for (int k = 0; k < 3; ++k)
			#pragma unroll
			for (int i = 0; i < 3; ++i)
				#pragma unroll
				for (int j = 0; j < 3; ++j) {
					real t = 0, u = 0;
					#pragma unroll
					for (int m = 0; m < 3; ++m) { t += U[iD + 6 * k + s6(i, m)] * gU[s6(m, j)]; u += gU[s6(k, m)] * U[iD + 6 * m + s6(i, j)]; }
					d_llu[k][i][j] = t; d_ull[k][i][j] = u;
				}
		// production form: d_ikl d_j^kl = d_il^m d_jm^l, so the doubly raised d_k^ij is never formed (18 fewer live values)
		if constexpr (!FAST) {
		#pragma unroll
		for (int k = 0; k < 3; ++k)
			#pragma unroll
			for (int i = 0; i < 3; ++i)
				#pragma unroll
				for (int j = 0; j < 3; ++j) {
					real t = 0;
					#pragma unroll
					for (int m = 0; m < 3; ++m) t += gU[s6(i, m)] * d_llu[k][m][j];
					d_luu[k][i][j] = t;
				}
		}
		real d_l[3], e_l[3];
		#pragma unroll
		for (int k = 0; k < 3; ++k) {
			d_l[k] = d_llu[k][0][0] + d_llu[k][1][1] + d_llu[k][2][2];
			e_l[k] = d_ull[0][0][k] + d_ull[1][1][k] + d_ull[2][2][k];
		}
		real const f_alpha = calc_f_alpha(s.f_eqn, alpha);
		real const f_alphaSq = calc_f_alphaSq(s.f_eqn, alpha);
		real const alphaSq_dalpha_f = calc_alphaSq_dalpha_f(s.f_eqn, alpha);
		deriv[iAlpha] += -f_alphaSq * trK;
		#pragma unroll
		for (int ij = 0; ij < 6; ++ij) deriv[iGamma + ij] += real(-2.) * alpha * K[ij];
		#pragma unroll
		for (int k = 0; k < 3; ++k) {
			real Kd = 0;
			#pragma unroll
			for (int i = 0; i < 3; ++i)
				#pragma unroll
				for (int j = 0; j < 3; ++j) Kd += K_uu[i][j] * U[iD + 6 * k + s6(i, j)];
			deriv[iA + k] += -(f_alpha + alphaSq_dalpha_f) * a[k] * trK + real(2.) * f_alpha * Kd;
			#pragma unroll
			for (int ij = 0; ij < 6; ++ij) deriv[iD + 6 * k + ij] += -alpha * a[k] * K[ij];
		}
		#pragma unroll
		for (int i = 0; i < 3; ++i)
			#pragma unroll
			for (int j = i; j < 3; ++j) {
				real t = -a[i] * a[j];
				#pragma unroll
				for (int k = 0; k < 3; ++k) {
					real const conn = d_llu[i][j][k] + d_llu[j][i][k] - d_ull[k][i][j];
					t += conn * (a[k] + d_l[k] - real(2.) * e_l[k]);
					#pragma unroll
					for (int l = 0; l < 3; ++l) {
						if constexpr (FAST) t += d_llu[i][l][k] * d_llu[j][k][l];
						else t += U[iD + 6 * i + s6(k, l)] * d_luu[j][k][l];
						t += real(2.) * d_llu[k][i][l] * (d_ull[k][j][l] - d_llu[l][j][k]);
					}
					t += real(-2.) * K[s6(i, k)] * K_ul[k][j];
				}
				t += trK * K[s6(i, j)];
				t += real(-8. * 3.14159265358979323846) * Sll[s6(i, j)] + real(4. * 3.14159265358979323846) * g[s6(i, j)] * (Str - rho);
				deriv[iK + s6(i, j)] += alpha * t;
			}
		#pragma unroll
		for (int k = 0; k < 3; ++k) deriv[iV + k] += (d_l[k] - e_l[k] - V[k]) * s.V_conv;
	}

	static HB_HD real calcDTCell(Params const& s, real const (&U)[nI], real const (&dx)[3], int dim) {
		real const f_alphaSq = calc_f_alphaSq(s.f_eqn, U[iAlpha]);
		const real* g = U + iGamma;
		real const det_gamma = det6(g);
		real const alpha_sqrt_f = rsqrt_ieee(f_alphaSq);
		real dt = inf_of<real>::v();
		if constexpr (FAST) {
			// production form: one reciprocal, one square root per side; dt = min_s dx_s / max(lambda_s, 1e-9)
			real const iDet = fastRcp(det_gamma);
			real const speed = rmax<real>(alpha_sqrt_f, U[iAlpha]);
			#pragma unroll
			for (int side = 0; side < 3; ++side) {
				if (side < dim && dx[side] > 0) {
					real const gjj = (side == 0 ? g[3] * g[5] - g[4] * g[4] : side == 1 ? g[0] * g[5] - g[2] * g[2] : g[0] * g[3] - g[1] * g[1]) * iDet;
					real y, sq;
					fastRsqrt(gjj, y, sq);
					dt = rmin<real>(dt, dx[side] * fastRcp(rmax<real>(real(1e-9), sq * speed)));
				}
			}
			return dt;
		} else {
		#pragma unroll
		for (int side = 0; side < 3; ++side) {
			if (side < dim && dx[side] > 0) {
				real gammaUjj;
				if (side == 0) gammaUjj = (g[3] * g[5] - g[4] * g[4]) / det_gamma;
				else if (side == 1) gammaUjj = (g[0] * g[5] - g[2] * g[2]) / det_gamma;
				else gammaUjj = (g[0] * g[3] - g[1] * g[1]) / det_gamma;
				real const sqrt_gammaUjj = rsqrt_ieee(gammaUjj);
				real const lambdaLight = sqrt_gammaUjj * U[iAlpha];
				real const lambdaGauge = sqrt_gammaUjj * alpha_sqrt_f;
				real const lambda = rmax<real>(lambdaGauge, lambdaLight);
				real const betaUi = 0.;
				real const lambdaMin = rmin<real>(real(0.), -betaUi - lambda);
				real const lambdaMax = rmax<real>(real(0.), -betaUi + lambda);
				real absLambdaMax = rmax<real>(rabs(lambdaMin), rabs(lambdaMax));
				absLambdaMax = rmax<real>(real(1e-9), absLambdaMax);
				dt = rmin<real>(dt, dx[side] / absLambdaMax);
			}
		}
		return dt;
		}
	}
	static HB_HD void constrainU(Params const&, real (&)[nI]) {}
	static HB_HD bool mirrorFlips(int, int) { return false; }   // mirror boundaries are rejected for this equation at hb_fv_create
};

}   // namespace hb
