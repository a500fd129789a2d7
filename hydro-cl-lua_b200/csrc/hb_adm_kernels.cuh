// hb_adm_kernels.cuh -- finite-volume stage of the ADM Bona-Masso 3-D system (BASELINE config C5) on sm_100a.
//
// One Runge-Kutta stage = the reference's calcFlux (Roe + flux limiter, hydro/solver/fvsolver.lua:57-198, hydro/flux/roe.cl:17-163),
// calcDerivFromFlux (hydro/solver/fvsolver.cl:6-125), addSource (hydro/eqn/adm3d.cl:1402-3072), the stage's multAdd combination
// (hydro/int/rk.lua:91-112) and, on the last stage, calcDT (hydro/eqn/cl/calcDT.cl:38-73), as two kernels over SoA arrays:
//   adm_flux<SIDE>  one thread per interface: the 13-wave Roe flux (a_side, d_side,ij, K_ij) -> flux scratch [side][13][cell]
//   adm_update      one thread per interior cell: flux differences of the 13 x dim components, source term of all 37, RK
//                   combination, store, CFL min (warp shuffle -> block -> atomicMin)
// The state is 51 reals per cell (408 B as an AoS record in the reference): only 20 of them enter a side's flux, so the flux
// kernel reads 4 cells x 20 variables per interface through L1/L2 and never touches the other 31.
// First version: not yet fused into one tile kernel like fv_march (next step for this equation).
#pragma once
#include "hb_fv_kernels.cuh"
#include "hb_eqn_adm3d.cuh"

namespace hb {

template<class Eqn, int SIDE>
HB_D void admLoadSide(typename Eqn::Side& s, const typename Eqn::real* __restrict__ U, long long idx, long long sv) {
	s.alpha = U[idx];
	#pragma unroll
	for (int k = 0; k < 6; ++k) s.g[k] = U[idx + (Eqn::iGamma + k) * sv];
	s.a = U[idx + (Eqn::iA + SIDE) * sv];
	#pragma unroll
	for (int k = 0; k < 6; ++k) s.d[k] = U[idx + (Eqn::iD + 6 * SIDE + k) * sv];
	#pragma unroll
	for (int k = 0; k < 6; ++k) s.K[k] = U[idx + (Eqn::iK + k) * sv];
}

// interfaces: the low face of cells c = g .. g+N along SIDE, interior in the transverse directions
template<class Eqn, int SIDE, int MODE>
__global__ void __launch_bounds__(128)
adm_flux(GridP<typename Eqn::real> const g, typename Eqn::Params const ep, const typename Eqn::real* __restrict__ U,
	typename Eqn::real* __restrict__ Fb, const double* dtPtr, int fluxLimiter)
{
	typedef typename Eqn::real real;
	int const n0 = g.N[0] + (SIDE == 0), n1 = g.N[1] + (SIDE == 1), n2 = g.N[2] + (SIDE == 2);
	long long const w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (w >= (long long)n0 * n1 * n2) return;
	int const i = int(w % n0) + HB_G;
	int const j = int((w / n0) % n1) + (g.dim >= 2 ? HB_G : 0);
	int const k = int(w / ((long long)n0 * n1)) + (g.dim >= 3 ? HB_G : 0);
	long long const idx = i + g.strideY * j + g.strideZ * k;
	long long const step = SIDE == 0 ? 1 : (SIDE == 1 ? g.strideY : g.strideZ);
	long long const sv = g.strideV;
	real Fa = 0, Fd[6], FK[6];
	#pragma unroll
	for (int q = 0; q < 6; ++q) { Fd[q] = 0; FK[q] = 0; }
	if (g.fluxOn[SIDE]) {
		typename Eqn::Side U2L, UL, UR, U2R;
		admLoadSide<Eqn, SIDE>(UL, U, idx - step, sv);
		admLoadSide<Eqn, SIDE>(UR, U, idx, sv);
		bool const lim = fluxLimiter > 0;
		if (lim) {
			admLoadSide<Eqn, SIDE>(U2L, U, idx - 2 * step, sv);
			admLoadSide<Eqn, SIDE>(U2R, U, idx + step, sv);
		} else { U2L = UL; U2R = UR; }
		real const dt_dx = real(*dtPtr) / g.dx[SIDE];
		Eqn::template roeFluxLimited<SIDE>(Fa, Fd, FK, ep, fluxLimiter, lim, dt_dx, U2L, UL, UR, U2R);
	}
	real* F = Fb + (long long)SIDE * 13 * sv + idx;
	F[0] = Fa;
	#pragma unroll
	for (int q = 0; q < 6; ++q) { F[(1 + q) * sv] = Fd[q]; F[(7 + q) * sv] = FK[q]; }
}

// The same flux with the characteristic differences shared between neighbouring interfaces: a thread computes the eigensystem and
// L (UR - UL) of ITS interface once and publishes the 13 differences in shared memory; the limiter of interface i reads those of
// i-1 and i+1 from there instead of rebuilding two more eigensystems from two more cells (3 -> 1 eigensystems, 4 -> 2 left
// transforms, 80 -> 40 loads per interface, ~half the registers).  The first and last thread along SIDE only publish.
// Thread layout: SIDE 0: NB threads along x;  SIDE 1, 2: 32 threads along x (coalescing) x NB interfaces along SIDE.
// Every published value is the same expression of the same cells as in adm_flux, so the strict build stays bit-identical.
// V: 1 = 32 x 8 threads, 2 blocks / SM (128 registers); 2 = 32 x 8, 1 block / SM; 3 = 16 x 8 threads (a 128-byte line per row), 3 blocks / SM;
//    5 = 8 x 16 threads (14 of 16 interfaces along SIDE productive instead of 6 of 8; 64-byte row segments), 3 blocks / SM
template<int SIDE, int V> struct AdmFluxGeom {
	static constexpr int LX = SIDE == 0 ? 1 : (V == 3 ? 16 : (V == 5 ? 8 : 32));     // threads along x (sides 1, 2)
	static constexpr int NB = SIDE == 0 ? 128 : (V == 5 ? 16 : 8);      // interfaces along SIDE per block, two of them halo
	static constexpr int NT = SIDE == 0 ? 128 : LX * NB;
	static constexpr int STEP = SIDE == 0 ? 1 : LX;                    // thread distance of the neighbouring interface
	static constexpr int MINB = SIDE == 0 ? 3 : (V == 1 ? 2 : (V == 2 ? 1 : 3));
};
template<class Eqn, int SIDE, int MODE, int V>
__global__ void __launch_bounds__((AdmFluxGeom<SIDE, V>::NT), (AdmFluxGeom<SIDE, V>::MINB))
adm_flux_shared(GridP<typename Eqn::real> const g, typename Eqn::Params const ep, const typename Eqn::real* __restrict__ U,
	typename Eqn::real* __restrict__ Fb, const double* dtPtr, int fluxLimiter)
{
	typedef typename Eqn::real real;
	typedef AdmFluxGeom<SIDE, V> G;
	constexpr int nW = Eqn::nW;
	__shared__ real sdU[nW][G::NT];
	int const tid = threadIdx.x;
	int const pos = SIDE == 0 ? tid : tid / G::LX;                               // position along SIDE inside the block
	int const gy = g.dim >= 2 ? HB_G : 0, gz = g.dim >= 3 ? HB_G : 0;
	int const c = HB_G - 1 + int(blockIdx.x) * (G::NB - 2) * (SIDE == 0) + int(blockIdx.y) * (G::NB - 2) * (SIDE != 0) + pos;   // interface = low face of cell c along SIDE
	int i, j, k;
	if (SIDE == 0) { i = c; j = int(blockIdx.y) + gy; k = int(blockIdx.z) + gz; }
	else if (SIDE == 1) { i = HB_G + int(blockIdx.x) * G::LX + tid % G::LX; j = c; k = int(blockIdx.z) + gz; }
	else { i = HB_G + int(blockIdx.x) * G::LX + tid % G::LX; j = int(blockIdx.z) + gy; k = c; }
	bool const active = c <= HB_G + g.N[SIDE] + 1 && (SIDE == 0 || i < HB_G + g.N[0]);
	long long const idx = i + g.strideY * j + g.strideZ * k;
	long long const step = SIDE == 0 ? 1 : (SIDE == 1 ? g.strideY : g.strideZ);
	long long const sv = g.strideV;
	typename Eqn::Eig eig;
	real fluxEig[nW], dUe[nW];
	if (active) {
		typename Eqn::Side UL, UR;
		admLoadSide<Eqn, SIDE>(UL, U, idx - step, sv);
		admLoadSide<Eqn, SIDE>(UR, U, idx, sv);
		Eqn::template interfaceEig<SIDE>(eig, ep, UL, UR);
		Eqn::template charAvg<SIDE>(fluxEig, eig, UL, UR);
		Eqn::template charDiff<SIDE>(dUe, eig, UL, UR);
		#pragma unroll
		for (int w = 0; w < nW; ++w) sdU[w][tid] = dUe[w];
	}
	__syncthreads();
	if (!active || pos == 0 || pos == G::NB - 1 || c > HB_G + g.N[SIDE]) return;
	real dUeL[nW], dUeR[nW];
	#pragma unroll
	for (int w = 0; w < nW; ++w) { dUeL[w] = sdU[w][tid - G::STEP]; dUeR[w] = sdU[w][tid + G::STEP]; }
	real Fa, Fd[6], FK[6];
	real const dt_dx = real(*dtPtr) / g.dx[SIDE];
	Eqn::template limitedFlux<SIDE>(Fa, Fd, FK, eig, fluxEig, dUe, dUeL, dUeR, fluxLimiter, true, dt_dx);
	real* F = Fb + (long long)SIDE * 13 * sv + idx;
	F[0] = Fa;
	#pragma unroll
	for (int q = 0; q < 6; ++q) { F[(1 + q) * sv] = Fd[q]; F[(7 + q) * sv] = FK[q]; }
}

// PART selects the integrated variables this launch produces (see inPart below); 2 = all 37 in one launch.
// The source of K_ij (adm3d.cl:1589-2776) needs the raised forms of all 18 d_kij at once (~75 live doubles); together with the
// other 31 derivatives and the RK combination of 37 variables one thread spills ~1 KB and the spill traffic evicts the state from L1:
// measured 2.0 ms per 128^3 launch against 0.27 ms for a 13-wave flux kernel.  Split in two launches every thread computes only
// what its part stores -- the arithmetic per stored value is the same expression sequence (the unused branches of the fully
// unrolled, compile-time indexed code are dead and dropped by the compiler), so the strict build stays bit-identical to the oracle.
template<class Eqn, int MODE, int PART>
__global__ void __launch_bounds__(128, PART == 0 ? 4 : (PART == 3 ? 4 : (PART >= 4 && PART <= 6 ? 8 : (PART == 7 ? 4 : (PART == 1 ? 3 : 1)))))
adm_update(GridP<typename Eqn::real> const g, StageP<typename Eqn::real> const sp, typename Eqn::Params const ep,
	const typename Eqn::real* __restrict__ Fb)
{
	typedef typename Eqn::real real;
	constexpr int nI = Eqn::nI;
	// PART 0: a_l, V_l (6); 1: K_ll (6); 7: alpha, gamma_ll (7, + the CFL minimum, which reads exactly these); 4, 5, 6: d_xij, d_yij, d_zij
	// (6 each); 3: all d_lll (18); 2: all 37.
	// The streaming parts (gamma_ll: -2 alpha K_ij; d_kij: flux difference - alpha a_k K_ij) are launched 6 variables at a time so that
	// they run at <= 64 registers: a memory-bound kernel needs occupancy, not a wide thread.
	auto inPart = [](int q) {
		bool const isK = q >= Eqn::iK && q < Eqn::iK + 6, isD = q >= Eqn::iD && q < Eqn::iK, isG = q == Eqn::iAlpha || (q >= Eqn::iGamma && q < Eqn::iGamma + 6);
		if (PART >= 4 && PART <= 6) return q >= Eqn::iD + 6 * (PART - 4) && q < Eqn::iD + 6 * (PART - 3);
		return PART == 2 || (PART == 1 && isK) || (PART == 3 && isD) || (PART == 7 && isG) || (PART == 0 && !isK && !isD && !isG);
	};
	__shared__ double redBuf[32];
	long long const nInt = (long long)g.N[0] * g.N[1] * g.N[2];
	long long const w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	real dtCell = inf_of<real>::v();
	if (w < nInt) {
		int const i = int(w % g.N[0]) + HB_G;
		int const j = int((w / g.N[0]) % g.N[1]) + (g.dim >= 2 ? HB_G : 0);
		int const k = int(w / ((long long)g.N[0] * g.N[1])) + (g.dim >= 3 ? HB_G : 0);
		long long const idx = i + g.strideY * j + g.strideZ * k;
		long long const sv = g.strideV;
		real const* __restrict__ Uin = sp.Uin;
		real U[nI], deriv[nI];
		#pragma unroll
		for (int q = 0; q < nI; ++q) { U[q] = Uin[idx + q * sv]; deriv[q] = 0; }
		if (sp.computeL) {
			// ---- calcDerivFromFlux (fvsolver.cl:97-123): sides in order; only the side's 13 components carry a flux
			if (g.volOn) {
				#pragma unroll
				for (int side = 0; side < 3; ++side) {
					if (side < g.dim) {
						long long const step = side == 0 ? 1 : (side == 1 ? g.strideY : g.strideZ);
						real const aov = g.aov[side];
						real const* FL = Fb + (long long)side * 13 * sv + idx;
						real const* FR = FL + step;
						deriv[Eqn::iA + side] = deriv[Eqn::iA + side] - (FR[0] * aov - FL[0] * aov);
						#pragma unroll
						for (int q = 0; q < 6; ++q) {
							deriv[Eqn::iD + 6 * side + q] = deriv[Eqn::iD + 6 * side + q] - (FR[(1 + q) * sv] * aov - FL[(1 + q) * sv] * aov);
							deriv[Eqn::iK + q] = deriv[Eqn::iK + q] - (FR[(7 + q) * sv] * aov - FL[(7 + q) * sv] * aov);
						}
					}
				}
			}
			// ---- addSource (adm3d.cl:1402-3072)
			real Sll[6];
			real const rho = Uin[idx + Eqn::iRho * sv];
			#pragma unroll
			for (int q = 0; q < 6; ++q) Sll[q] = Uin[idx + (Eqn::iSll + q) * sv];
			Eqn::addSource(deriv, ep, U, rho, Sll);
			if (ep.a_conv != 0 || ep.d_conv != 0) {
				// first-order constraint convergence (adm3d.cl:3008-3036): radius-1 centred differences of alpha and gamma_ll
				#pragma unroll
				for (int side = 0; side < 3; ++side) {
					if (side < g.dim) {
						long long const step = side == 0 ? 1 : (side == 1 ? g.strideY : g.strideZ);
						real const dx = g.dx[side];
						real const partial_i_log_alpha = (log(Uin[idx + step]) - log(Uin[idx - step])) / (real(2.) * dx);
						deriv[Eqn::iA + side] += ep.a_conv * (partial_i_log_alpha - U[Eqn::iA + side]);
						#pragma unroll
						for (int q = 0; q < 6; ++q) {
							real const pg = (Uin[idx + step + (Eqn::iGamma + q) * sv] - Uin[idx - step + (Eqn::iGamma + q) * sv]) / (real(2.) * dx);
							deriv[Eqn::iD + 6 * side + q] += ep.d_conv * (real(.5) * pg - U[Eqn::iD + 6 * side + q]);
						}
					} else {
						deriv[Eqn::iA + side] += ep.a_conv * (real(0.) - U[Eqn::iA + side]);
						#pragma unroll
						for (int q = 0; q < 6; ++q) deriv[Eqn::iD + 6 * side + q] += ep.d_conv * (real(.5) * real(0) - U[Eqn::iD + 6 * side + q]);
					}
				}
			}
		}
		if (sp.Lout) {
			#pragma unroll
			for (int q = 0; q < nI; ++q) if (inPart(q)) sp.Lout[idx + q * sv] = deriv[q];
		}
		if (sp.Uout) {
			double const dt = *sp.dt;
			#pragma unroll
			for (int q = 0; q < nI; ++q) if (inPart(q)) {
				real r = 0;
				#pragma unroll
				for (int a = 0; a < HB_MAX_TERMS; ++a)
					if (a < sp.nA) r = r + (((sp.aOwnMask >> a) & 1) ? U[q] : sp.aPtr[a][idx + q * sv]) * real(sp.aCoef[a]);
				#pragma unroll
				for (int b = 0; b < HB_MAX_TERMS; ++b)
					if (b < sp.nB) r = r + sp.bPtr[b][idx + q * sv] * real(sp.bCoef[b] * dt);
				if (sp.computeL) r = r + deriv[q] * real(sp.betaSelf * dt);
				deriv[q] = r;          // reuse as the new state
			}
			#pragma unroll
			for (int q = 0; q < nI; ++q) if (inPart(q)) sp.Uout[idx + q * sv] = deriv[q];
			if ((PART == 7 || PART == 2) && sp.dtMinBits) dtCell = Eqn::calcDTCell(ep, deriv, g.dx, g.dim);   // reads alpha, gamma_ll only
		}
	}
	if ((PART == 7 || PART == 2) && sp.dtMinBits) {
		double v = double(dtCell);
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) { double const u = __shfl_xor_sync(0xffffffffu, v, o); v = u < v ? u : v; }
		if ((threadIdx.x & 31) == 0) redBuf[threadIdx.x >> 5] = v;
		__syncthreads();
		if (threadIdx.x < 32) {
			v = threadIdx.x < (blockDim.x + 31) / 32 ? redBuf[threadIdx.x] : HUGE_VAL;
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) { double const u = __shfl_xor_sync(0xffffffffu, v, o); v = u < v ? u : v; }
			if (threadIdx.x == 0 && v < HUGE_VAL) atomicMin(sp.dtMinBits, dtBits(v));
		}
	}
}

// initDerivs (hydro/eqn/adm3d.cl:196-243): a_l, d_lll from centred differences of alpha, gamma_ll (interior cells), V_i = d_ik^k - d^k_ki.
// Reads alpha / gamma_ll (never written here) of the neighbours, writes a_l, d_lll, V_l of its own cell: safe in place.
template<class Eqn, int MODE>
__global__ void adm_init_derivs(GridP<typename Eqn::real> const g, typename Eqn::real* __restrict__ U)
{
	typedef typename Eqn::real real;
	long long const nInt = (long long)g.N[0] * g.N[1] * g.N[2];
	long long const w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (w >= nInt) return;
	int const i = int(w % g.N[0]) + HB_G;
	int const j = int((w / g.N[0]) % g.N[1]) + (g.dim >= 2 ? HB_G : 0);
	int const k = int(w / ((long long)g.N[0] * g.N[1])) + (g.dim >= 3 ? HB_G : 0);
	long long const idx = i + g.strideY * j + g.strideZ * k;
	long long const sv = g.strideV;
	real gl[6], gU[6], d[3][6], a[3];
	#pragma unroll
	for (int q = 0; q < 6; ++q) gl[q] = U[idx + (Eqn::iGamma + q) * sv];
	Eqn::inv6(gU, gl, Eqn::det6(gl));
	real const alpha = U[idx];
	#pragma unroll
	for (int side = 0; side < 3; ++side) {
		if (side < g.dim) {
			long long const step = side == 0 ? 1 : (side == 1 ? g.strideY : g.strideZ);
			real const dx = g.dx[side];
			a[side] = (U[idx + step] - U[idx - step]) / (dx * alpha);
			#pragma unroll
			for (int q = 0; q < 6; ++q)
				d[side][q] = real(.5) * (U[idx + step + (Eqn::iGamma + q) * sv] - U[idx - step + (Eqn::iGamma + q) * sv]) / dx;
		} else {
			a[side] = 0;
			#pragma unroll
			for (int q = 0; q < 6; ++q) d[side][q] = 0;
		}
	}
	#pragma unroll
	for (int s = 0; s < 3; ++s) {
		U[idx + (Eqn::iA + s) * sv] = a[s];
		#pragma unroll
		for (int q = 0; q < 6; ++q) U[idx + (Eqn::iD + 6 * s + q) * sv] = d[s][q];
	}
	#pragma unroll
	for (int ii = 0; ii < 3; ++ii) {
		real t = 0.;
		#pragma unroll
		for (int jj = 0; jj < 3; ++jj)
			#pragma unroll
			for (int kk = 0; kk < 3; ++kk)
				t = t + gU[Eqn::s6(jj, kk)] * (d[ii][Eqn::s6(jj, kk)] - d[jj][Eqn::s6(kk, ii)]);
		U[idx + (Eqn::iV + ii) * sv] = t;
	}
}

}   // namespace hb
