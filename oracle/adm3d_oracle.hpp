// adm3d_oracle.hpp -- TEST INFRASTRUCTURE ONLY (included by hydro_oracle.cpp).
//
// CPU restatement of the ADM Bona-Masso 3-D equation plug-in of the reference (hydro/eqn/adm3d.lua, adm3d.cl,
// hydro/eqn/einstein.lua, hydro/init/einstein.lua) as the reference runs it for BASELINE config C5:
// noZeroRowsInFlux = true (13 waves), useShift = 'none', roeUseFluxFromCons = false, 'constrain V' = none.
//
// State (hydro/eqn/adm3d.lua:65-75,144-163), 51 reals: alpha, gamma_ll[6], a_l[3], d_lll[3][6], K_ll[6], V_l[3]   (37 integrated)
//                                                     rho, S_u[3], S_ll[6], H, M_u[3]                              (14 not integrated)
// Symmetric 3x3 storage order xx, xy, xz, yy, yz, zz (hydro/code/math.cl real3s3).
//
// PARITY PIN STATUS.  eigen_forInterface / eigen_leftTransform / eigen_rightTransform / wave speeds / initDerivs / calcDT follow the
// reference expression by expression (file:line cited at each function).  The source term is different: the reference's
// addSource is a 1190-line machine-generated sequence of `tmpN` products (adm3d.cl:1589-2776); it is restated here in tensor
// form, obtained by fitting a basis of tensor contractions to that polynomial symbolically (tools/adm_source_fit.py, exact
// rational coefficients, unique solution):
//     d/dt alpha    += -alpha^2 f K
//     d/dt gamma_ij += -2 alpha K_ij
//     d/dt a_k      += -(alpha f + alpha^2 f') a_k K + 2 alpha f K^ij d_kij
//     d/dt d_kij    += -alpha a_k K_ij
//     d/dt K_ij     += alpha ( -a_i a_j + conn^k_ij (a_k + d_k - 2 e_k) + d_ikl d_j^kl + 2 d_kil d^k_j^l - 2 d_kil d^l_j^k
//                              + K K_ij - 2 K_ik K^k_j - 8 pi S_ij + 4 pi gamma_ij (S - rho) )
//     d/dt V_k      += V_convCoeff (d_k - e_k - V_k)                                           (adm3d.cl:3060-3068)
//   with conn^k_ij = d_ij^k + d_ji^k - d^k_ij, d_k = d_km^m, e_k = d^m_mk, K = K^m_m.
// The value of the two forms differs by summation order only; tests/test_adm3d.py checks the tensor form against the reference's
// own generated lines compiled into oracle/_ref/ (oracle/build_ref.py) to 1e-13.
// Symmath-emitted gauge helpers calc_f* (hydro/init/einstein.lua:54-81) have unknown association order: "parity unpinned".
#pragma once

namespace ho {

template<class real_> struct ADM3D {
	typedef real_ real;
	enum { numStates = 51, numIntStates = 37, numWaves = 13 };
	static constexpr bool roeUseFluxFromCons = false;   // adm3d.lua:21
	static constexpr bool hasEigenForCell = false;
	static constexpr bool isEuler = false;
	static constexpr bool hasWaveMinMax = false;        // hll / rusanov are not restated for this equation
	static constexpr bool hasSource = true;
	enum { iAlpha = 0, iGamma = 1, iA = 7, iD = 10, iK = 28, iV = 34, iRho = 37, iSu = 38, iSll = 41, iH = 47, iMu = 48 };
	struct cons_t { real ptr[numStates]; };
	struct waves_t { real ptr[numWaves]; };
	struct eigen_t { real alpha, alpha_sqrt_f; real gamma_uu[6]; real sqrt_gammaUjj[3]; };   // adm3d.lua:165-171

	static inline int s6(int i, int j) { static const int t[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}}; return t[i][j]; }

	// ---- gauge function f(alpha) and the products the reference pre-computes (hydro/init/einstein.lua:54-81;
	//      options hydro/eqn/einstein.lua:42-48)
	static real fConst(int id) { static const double c[8] = {0, 0, 1., 0., .49, .5, 1.5, 1.69}; return real(c[id]); }
	static real calc_f(int id, real a) { return id == 0 ? real(2.) / a : id == 1 ? real(1.) + real(1.) / (a * a) : fConst(id); }
	static real calc_f_alpha(int id, real a) { return id == 0 ? real(2.) : id == 1 ? a + real(1.) / a : fConst(id) * a; }
	static real calc_f_alphaSq(int id, real a) { return id == 0 ? real(2.) * a : id == 1 ? a * a + real(1.) : fConst(id) * a * a; }
	static real calc_alphaSq_dalpha_f(int id, real a) { return id == 0 ? real(-2.) : id == 1 ? real(-2.) / a : real(0.); }

	// ---- real3s3 helpers (hydro/code/math.cl:518-536,591-594,683-685)
	static real det6(const real* m) {
		return m[0] * (m[3] * m[5] - m[4] * m[4]) - m[1] * (m[1] * m[5] - m[2] * m[4]) + m[2] * (m[1] * m[4] - m[3] * m[2]);
	}
	static void inv6(real* o, const real* m, real det) {
		real const invDet = real(1.) / det;
		o[0] = (m[3] * m[5] - m[4] * m[4]) * invDet;
		o[1] = (m[2] * m[4] - m[1] * m[5]) * invDet;
		o[2] = (m[1] * m[4] - m[2] * m[3]) * invDet;
		o[3] = (m[0] * m[5] - m[2] * m[2]) * invDet;
		o[4] = (m[2] * m[1] - m[0] * m[4]) * invDet;
		o[5] = (m[0] * m[3] - m[1] * m[1]) * invDet;
	}
	static real dot6(const real* a, const real* b) {
		return a[0] * b[0] + a[3] * b[3] + a[5] * b[5] + real(2.) * (a[1] * b[1] + a[2] * b[2] + a[4] * b[4]);
	}
	static void swap6(real* o, const real* m, int side) {
		static const int p[3][6] = {{0, 1, 2, 3, 4, 5}, {3, 1, 4, 0, 2, 5}, {5, 4, 2, 3, 1, 0}};
		for (int k = 0; k < 6; ++k) o[k] = m[p[side][k]];
	}

	// adm3d.cl:373-397
	template<class S> static void eigen_forInterface(eigen_t& eig, S const& solver, cons_t const& UL, cons_t const& UR, normal_t) {
		eig.alpha = real(.5) * (UL.ptr[iAlpha] + UR.ptr[iAlpha]);
		real avg[6];
		for (int k = 0; k < 6; ++k) avg[k] = (UL.ptr[iGamma + k] + UR.ptr[iGamma + k]) * real(.5);
		real const det = det6(avg);
		eig.alpha_sqrt_f = std::sqrt(calc_f_alphaSq(solver.f_eqn, eig.alpha));
		inv6(eig.gamma_uu, avg, det);
		eig.sqrt_gammaUjj[0] = std::sqrt(eig.gamma_uu[0]);
		eig.sqrt_gammaUjj[1] = std::sqrt(eig.gamma_uu[3]);
		eig.sqrt_gammaUjj[2] = std::sqrt(eig.gamma_uu[5]);
	}
	// adm3d.lua:427-479 (noZeroRowsInFlux, no shift: betaUi = 0)
	template<class S> static void eigenWaves(real* lambdas, S const&, eigen_t const& eig, normal_t n) {
		real const sq = eig.sqrt_gammaUjj[n.side];
		real const lambdaLight = sq * eig.alpha;
		real const lambdaGauge = sq * eig.alpha_sqrt_f;
		lambdas[0] = -real(0) - lambdaGauge;
		for (int k = 1; k <= 5; ++k) lambdas[k] = -real(0) - lambdaLight;
		lambdas[6] = -real(0);
		for (int k = 7; k <= 11; ++k) lambdas[k] = -real(0) + lambdaLight;
		lambdas[12] = -real(0) + lambdaGauge;
	}
	// adm3d.cl:670-707
	template<class S> static void eigen_leftTransform(waves_t& result, S const&, eigen_t const& eig, cons_t const& inputU, normal_t n) {
		real const _1_sqrt_f = eig.alpha / eig.alpha_sqrt_f;
		real const _1_f = _1_sqrt_f * _1_sqrt_f;
		int const side = n.side;
		real const sqrt_gammaUjj = eig.sqrt_gammaUjj[side];
		real const _1_gammaUjj = real(1.) / eig.gamma_uu[s6(side, side)];
		real const a_j = inputU.ptr[iA + side];
		real d_lll[6], K_ll[6], gamma_uu[6];
		swap6(d_lll, inputU.ptr + iD + 6 * side, side);
		swap6(K_ll, inputU.ptr + iK, side);
		swap6(gamma_uu, eig.gamma_uu, side);
		real const K_dot_eig_gamma = dot6(K_ll, gamma_uu);
		real const dj_dot_eig_gamma = dot6(d_lll, gamma_uu);
		result.ptr[0] = (a_j * -sqrt_gammaUjj * _1_sqrt_f + K_dot_eig_gamma) * real(.5) * _1_gammaUjj;
		for (int i = 1; i <= 5; ++i) result.ptr[i] = real(.5) * (-sqrt_gammaUjj * d_lll[i] + K_ll[i]);
		result.ptr[6] = (-a_j * _1_f + dj_dot_eig_gamma) * _1_gammaUjj;
		for (int i = 1; i <= 5; ++i) result.ptr[6 + i] = real(.5) * (sqrt_gammaUjj * d_lll[i] + K_ll[i]);
		result.ptr[12] = (a_j * sqrt_gammaUjj * _1_sqrt_f + K_dot_eig_gamma) * real(.5) * _1_gammaUjj;
	}
	// adm3d.cl:712-722,1225-1274
	template<class S> static void eigen_rightTransform(cons_t& result, S const&, eigen_t const& eig, waves_t const& input, normal_t n) {
		for (int j = 0; j < numStates; ++j) result.ptr[j] = 0;
		int const side = n.side;
		real gamma_uu[6];
		swap6(gamma_uu, eig.gamma_uu, side);
		real const input1_dot_gammaU = input.ptr[1] * real(2.) * gamma_uu[1]
			+ input.ptr[2] * real(2.) * gamma_uu[2]
			+ input.ptr[3] * gamma_uu[3]
			+ input.ptr[4] * real(2.) * gamma_uu[4]
			+ input.ptr[5] * gamma_uu[5];
		real const input7_dot_gammaU = input.ptr[7] * real(2.) * gamma_uu[1]
			+ input.ptr[8] * real(2.) * gamma_uu[2]
			+ input.ptr[9] * gamma_uu[3]
			+ input.ptr[10] * real(2.) * gamma_uu[4]
			+ input.ptr[11] * gamma_uu[5];
		real const sqrt_f = eig.alpha_sqrt_f / eig.alpha;
		real const _1_sqrt_f = real(1.) / sqrt_f;
		real const sqrt_gammaUjj = eig.sqrt_gammaUjj[side];
		real const _1_sqrt_gammaUjj = real(1.) / sqrt_gammaUjj;
		real const _1_gammaUjj = _1_sqrt_gammaUjj * _1_sqrt_gammaUjj;
		result.ptr[iA + side] = sqrt_f * sqrt_gammaUjj * (input.ptr[12] - input.ptr[0]);
		real d_lll[6], K_ll[6];
		d_lll[0] = ((input.ptr[12] - input.ptr[0]) * _1_sqrt_f + (input1_dot_gammaU - input7_dot_gammaU) * _1_gammaUjj) * _1_sqrt_gammaUjj + input.ptr[6];
		for (int i = 1; i <= 5; ++i) d_lll[i] = (input.ptr[i + 6] - input.ptr[i]) * _1_sqrt_gammaUjj;
		K_ll[0] = input.ptr[0] + input.ptr[12] - (input1_dot_gammaU + input7_dot_gammaU) * _1_gammaUjj;
		for (int i = 1; i <= 5; ++i) K_ll[i] = input.ptr[i] + input.ptr[i + 6];
		swap6(result.ptr + iD + 6 * side, d_lll, side);
		swap6(result.ptr + iK, K_ll, side);
	}
	template<class S> static void fluxFromCons(cons_t&, S const&, cons_t const&, normal_t) {}   // not used: roeUseFluxFromCons = false

	// eqn.lua:1187-1224 with adm3d.lua:500-544 (consWaveCodeMinMaxAllSides)
	template<class S> static void calcDTCell(real& dt, S const& solver, cons_t const& U) {
		real const f_alphaSq = calc_f_alphaSq(solver.f_eqn, U.ptr[iAlpha]);
		const real* g = U.ptr + iGamma;
		real const det_gamma = det6(g);
		real const alpha_sqrt_f = std::sqrt(f_alphaSq);
		for (int side = 0; side < solver.dim; ++side) {
			real const dx = solver.grid_dx.s(side);
			if (dx > 0) {
				real gammaUjj;
				if (side == 0) gammaUjj = (g[3] * g[5] - g[4] * g[4]) / det_gamma;
				else if (side == 1) gammaUjj = (g[0] * g[5] - g[2] * g[2]) / det_gamma;
				else gammaUjj = (g[0] * g[3] - g[1] * g[1]) / det_gamma;
				real const sqrt_gammaUjj = std::sqrt(gammaUjj);
				real const lambdaLight = sqrt_gammaUjj * U.ptr[iAlpha];
				real const lambdaGauge = sqrt_gammaUjj * alpha_sqrt_f;
				real const lambda = clmax<real>(lambdaGauge, lambdaLight);
				real const betaUi = 0.;
				real const lambdaMin = clmin<real>(real(0.), -betaUi - lambda);
				real const lambdaMax = clmax<real>(real(0.), -betaUi + lambda);
				real absLambdaMax = clmax<real>(std::fabs(lambdaMin), std::fabs(lambdaMax));
				absLambdaMax = clmax<real>(real(1e-9), absLambdaMax);
				dt = clmin<real>(dt, dx / absLambdaMax);
			}
		}
	}

	// constrainU (adm3d.cl:3077-3290) with 'constrain V' = none only recomputes the diagnostic fields H and M_u, which are not
	// integrated and do not feed back; they are left at zero here and in the CUDA path (out of scope)
	template<class S> static void constrainU(S const&, cons_t&) {}
	static void mirrorReflect(cons_t&, int) {}   // mirror boundaries are not used with this equation in the configs

	// ---- addSource (adm3d.cl:1402-3072), tensor form (see the header).  U = this cell; the constraint-convergence terms with
	// a_convCoeff / d_convCoeff (radius-1 stencil, default 0) are applied by the caller-supplied neighbour accessor only when non-zero.
	template<class S, class NB> static void addSource(cons_t& deriv, S const& solver, cons_t const& U, NB&& nb) {
		real const alpha = U.ptr[iAlpha];
		const real* g = U.ptr + iGamma; const real* a = U.ptr + iA; const real* K = U.ptr + iK; const real* V = U.ptr + iV;
		real gU[6];
		inv6(gU, g, det6(g));
		auto d = [&](int k, int i, int j) { return U.ptr[iD + 6 * k + s6(i, j)]; };
		auto GU = [&](int i, int j) { return gU[s6(i, j)]; };
		real const rho = U.ptr[iRho];
		real const Str = dot6(U.ptr + iSll, gU);
		// K^i_j, K, K^ij
		real K_ul[3][3], K_uu[3][3];
		for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { real s = 0; for (int m = 0; m < 3; ++m) s += GU(i, m) * K[s6(m, j)]; K_ul[i][j] = s; }
		for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { real s = 0; for (int m = 0; m < 3; ++m) s += K_ul[i][m] * GU(m, j); K_uu[i][j] = s; }
		real const trK = K_ul[0][0] + K_ul[1][1] + K_ul[2][2];
		// d_ki^j, d^k_ij, d_k^ij
		real d_llu[3][3][3], d_ull[3][3][3], d_luu[3][3][3];
		for (int k = 0; k < 3; ++k) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
			real s = 0, t = 0;
			for (int m = 0; m < 3; ++m) { s += d(k, i, m) * GU(m, j); t += GU(k, m) * d(m, i, j); }
			d_llu[k][i][j] = s; d_ull[k][i][j] = t;
		}
		for (int k = 0; k < 3; ++k) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
			real s = 0;
			for (int m = 0; m < 3; ++m) s += GU(i, m) * d_llu[k][m][j];
			d_luu[k][i][j] = s;
		}
		real d_l[3], e_l[3];
		for (int k = 0; k < 3; ++k) {
			d_l[k] = d_llu[k][0][0] + d_llu[k][1][1] + d_llu[k][2][2];
			e_l[k] = d_ull[0][0][k] + d_ull[1][1][k] + d_ull[2][2][k];
		}
		real const f_alpha = calc_f_alpha(solver.f_eqn, alpha);
		real const f_alphaSq = calc_f_alphaSq(solver.f_eqn, alpha);
		real const alphaSq_dalpha_f = calc_alphaSq_dalpha_f(solver.f_eqn, alpha);
		deriv.ptr[iAlpha] += -f_alphaSq * trK;
		for (int ij = 0; ij < 6; ++ij) deriv.ptr[iGamma + ij] += real(-2.) * alpha * K[ij];
		for (int k = 0; k < 3; ++k) {
			real Kd = 0;
			for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Kd += K_uu[i][j] * d(k, i, j);
			deriv.ptr[iA + k] += -(f_alpha + alphaSq_dalpha_f) * a[k] * trK + real(2.) * f_alpha * Kd;
			for (int ij = 0; ij < 6; ++ij) deriv.ptr[iD + 6 * k + ij] += -alpha * a[k] * K[ij];
		}
		for (int i = 0; i < 3; ++i) for (int j = i; j < 3; ++j) {
			real s = -a[i] * a[j];
			for (int k = 0; k < 3; ++k) {
				real const conn = d_llu[i][j][k] + d_llu[j][i][k] - d_ull[k][i][j];      // conn^k_ij
				s += conn * (a[k] + d_l[k] - real(2.) * e_l[k]);
				for (int l = 0; l < 3; ++l) {
					s += d(i, k, l) * d_luu[j][k][l];                                    // d_ikl d_j^kl
					s += real(2.) * d_llu[k][i][l] * (d_ull[k][j][l] - d_llu[l][j][k]);   // 2 d_ki^l (d^k_jl - d_lj^k)
				}
				s += real(-2.) * K[s6(i, k)] * K_ul[k][j];
			}
			s += trK * K[s6(i, j)];
			s += real(-8. * M_PI) * U.ptr[iSll + s6(i, j)] + real(4. * M_PI) * g[s6(i, j)] * (Str - rho);
			deriv.ptr[iK + s6(i, j)] += alpha * s;
		}
		for (int k = 0; k < 3; ++k) deriv.ptr[iV + k] += (d_l[k] - e_l[k] - V[k]) * solver.V_convCoeff;
		// first-order constraint convergence (adm3d.cl:3008-3036), off by default
		if (solver.a_convCoeff != 0 || solver.d_convCoeff != 0) {
			for (int i = 0; i < solver.dim; ++i) {
				cons_t const& Up = nb(i, +1); cons_t const& Um = nb(i, -1);
				real const dx = solver.grid_dx.s(i);
				real const partial_i_log_alpha = (std::log(Up.ptr[iAlpha]) - std::log(Um.ptr[iAlpha])) / (real(2.) * dx);
				deriv.ptr[iA + i] += solver.a_convCoeff * (partial_i_log_alpha - a[i]);
				for (int jk = 0; jk < 6; ++jk) {
					real const partial_i_gamma_jk = (Up.ptr[iGamma + jk] - Um.ptr[iGamma + jk]) / (real(2.) * dx);
					deriv.ptr[iD + 6 * i + jk] += solver.d_convCoeff * (real(.5) * partial_i_gamma_jk - U.ptr[iD + 6 * i + jk]);
				}
			}
			for (int i = solver.dim; i < 3; ++i) {
				deriv.ptr[iA + i] += solver.a_convCoeff * (real(0.) - a[i]);
				for (int jk = 0; jk < 6; ++jk) deriv.ptr[iD + 6 * i + jk] += solver.d_convCoeff * (real(.5) * real(0) - U.ptr[iD + 6 * i + jk]);
			}
		}
	}

	// ---- initDerivs (adm3d.cl:196-243): a_l, d_lll by centred differences of alpha, gamma_ll; V_i = d_ik^k - d^k_ki
	template<class S, class NB> static void initDerivs(cons_t& U, S const& solver, NB&& nb) {
		real gU[6];
		inv6(gU, U.ptr + iGamma, det6(U.ptr + iGamma));
		for (int i = 0; i < solver.dim; ++i) {
			cons_t const& Up = nb(i, +1); cons_t const& Um = nb(i, -1);
			real const dx = solver.grid_dx.s(i);
			U.ptr[iA + i] = (Up.ptr[iAlpha] - Um.ptr[iAlpha]) / (dx * U.ptr[iAlpha]);
			for (int jk = 0; jk < 6; ++jk) U.ptr[iD + 6 * i + jk] = real(.5) * (Up.ptr[iGamma + jk] - Um.ptr[iGamma + jk]) / dx;
		}
		for (int i = solver.dim; i < 3; ++i) {
			U.ptr[iA + i] = 0;
			for (int jk = 0; jk < 6; ++jk) U.ptr[iD + 6 * i + jk] = 0;
		}
		for (int i = 0; i < 3; ++i) {
			real s = 0.;
			for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k)
				s = s + gU[s6(j, k)] * (U.ptr[iD + 6 * i + s6(j, k)] - U.ptr[iD + 6 * j + s6(k, i)]);
			U.ptr[iV + i] = s;
		}
	}
};

}   // namespace ho
