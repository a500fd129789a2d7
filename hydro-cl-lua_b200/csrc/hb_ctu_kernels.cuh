// hb_ctu_kernels.cuh -- the corner-transport-upwind variant of FiniteVolumeSolver:calcDeriv:
//   hydro/solver/fvsolver.lua:225-302 with useCTU: calcLR, calcFlux, updateCTU (hydro/solver/ctu.cl:12-129), boundaryLR
//   (hydro/solver/gridsolver.lua:463-473,1241-1268), calcFlux again, calcDerivFromFlux (hydro/solver/fvsolver.cl:6-125).
//
// updateCTU advances BOTH face states of EVERY side of a cell by half a step of the flux differences of ALL sides, and the second flux
// pass reads the corrected face states of the neighbours: the data flow between the two flux passes spans the whole stencil twice, which
// is exactly what the fused stage kernels avoid materialising.  This variant therefore keeps the reference's buffers -- ULR (2 x dim
// face-state records per cell) and flux (dim records per cell), as structure-of-arrays blocks of the solver's layout -- and its kernel
// sequence, one thread per cell, every access coalesced along x.  HBM traffic per cell and stage in words of nI reals (3-D):
// calcLR 1 + 6, calcFlux 2 x (6 + 3), updateCTU 6 + 3 + 6, boundaryLR ~0, finish 6 + RK operands: ~45, against ~4 of the fused kernel.
#pragma once
#include "hb_fv_kernels.cuh"

namespace hb {

template<class real> struct CtuP {
	real* ULR;               // block b = 2 * side + (0: L, 1: R), each laid out like a state of nI variables (variable stride g.strideV)
	real* flux;              // block b = side
	long long blockStride;   // elements between blocks
	real areaL[3], areaR[3], invVolume;   // ctu.cl:26-62 on a cartesian grid (weightFluxByGridVolume, cell->volume = prod grid_dx), formed on the host in `real`
};

constexpr int HB_CTU_NT = 128;

template<class real> HB_D bool ctuCell(GridP<real> const& g, int& i, int& j, int& k, long long& idx) {
	long long const w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	long long const S0 = g.S[0], S1 = g.S[1], S2 = g.S[2];
	if (w >= S0 * S1 * S2) return false;
	i = int(w % S0); j = int((w / S0) % S1); k = int(w / (S0 * S1));
	idx = i + g.strideY * j + g.strideZ * k;
	return true;
}
template<class real> HB_D bool ctuOOB(GridP<real> const& g, int i, int j, int k, int l, int r) {
	return i < l || i >= g.S[0] - r || (g.dim >= 2 && (j < l || j >= g.S[1] - r)) || (g.dim >= 3 && (k < l || k >= g.S[2] - r));
}
template<class real> HB_D long long ctuStride(GridP<real> const& g, int s) { return s == 0 ? 1 : (s == 1 ? g.strideY : g.strideZ); }

// calcLR, 'plm cons' (plm.cl:32-91,976-997; SETBOUNDS(1,1))
template<class Eqn, int MODE>
__global__ void ctu_calc_lr(GridP<typename Eqn::real> const g, StageP<typename Eqn::real> const sp, CtuP<typename Eqn::real> const c)
{
	typedef typename Eqn::real real;
	constexpr int nI = Eqn::nI;
	int i, j, k; long long idx;
	if (!ctuCell(g, i, j, k, idx) || ctuOOB(g, i, j, k, 1, 1)) return;
	for (int side = 0; side < g.dim; ++side) {
		long long const st = ctuStride(g, side);
		real* L = c.ULR + (2 * side) * c.blockStride + idx;
		real* R = c.ULR + (2 * side + 1) * c.blockStride + idx;
		#pragma unroll
		for (int q = 0; q < nI; ++q) {
			real const* u = sp.Uin + q * g.strideV + idx;
			real const U = u[0];
			real const h = plmHalfSlope<real>(sp.slopeLimiter, u[-st], U, u[st]);
			R[q * g.strideV] = U + h;
			L[q * g.strideV] = U - h;
		}
	}
}

template<class Eqn, int SIDE>
HB_D void ctuFluxSide(GridP<typename Eqn::real> const& g, StageP<typename Eqn::real> const& sp, CtuP<typename Eqn::real> const& c,
	typename Eqn::Params const& ep, long long idx)
{
	typedef typename Eqn::real real;
	constexpr int nI = Eqn::nI;
	real F[nI];
	if (!g.fluxOn[SIDE]) {                       // fvsolver.lua:107-110: area <= 1e-7
		#pragma unroll
		for (int q = 0; q < nI; ++q) F[q] = 0;
	} else {
		long long const st = ctuStride(g, SIDE);
		real const* Lr = c.ULR + (2 * SIDE + 1) * c.blockStride + (idx - st);   // ULR[side + dim * indexL].R
		real const* Rl = c.ULR + (2 * SIDE) * c.blockStride + idx;              // ULR[side + dim * indexR].L
		real UL[nI], UR[nI];
		#pragma unroll
		for (int q = 0; q < nI; ++q) { UL[q] = Lr[q * g.strideV]; UR[q] = Rl[q * g.strideV]; }
		interfaceFlux<Eqn, SIDE>(sp.flux, sp.fluxParam, F, ep, UL, UR);
	}
	real* out = c.flux + SIDE * c.blockStride + idx;
	#pragma unroll
	for (int q = 0; q < nI; ++q) out[q * g.strideV] = F[q];
}

// calcFlux on the stored face states (fvsolver.lua:57-198 with usePLM: gridsolver.lua:486-496; OOB(numGhost, numGhost - 1))
template<class Eqn, int MODE>
__global__ void ctu_calc_flux(GridP<typename Eqn::real> const g, StageP<typename Eqn::real> const sp, CtuP<typename Eqn::real> const c,
	typename Eqn::Params const ep)
{
	int i, j, k; long long idx;
	if (!ctuCell(g, i, j, k, idx) || ctuOOB(g, i, j, k, HB_G, HB_G - 1)) return;
	ctuFluxSide<Eqn, 0>(g, sp, c, ep, idx);
	if (g.dim >= 2) ctuFluxSide<Eqn, 1>(g, sp, c, ep, idx);
	if (g.dim >= 3) ctuFluxSide<Eqn, 2>(g, sp, c, ep, idx);
}

template<class Eqn, int SIDE>
HB_D void ctuUpdateSide(GridP<typename Eqn::real> const& g, CtuP<typename Eqn::real> const& c, typename Eqn::Params const& ep,
	long long idx, typename Eqn::real dt)
{
	typedef typename Eqn::real real;
	constexpr int nI = Eqn::nI;
	real* Lp = c.ULR + (2 * SIDE) * c.blockStride + idx;
	real* Rp = c.ULR + (2 * SIDE + 1) * c.blockStride + idx;
	real UL[nI], UR[nI], fluxCellL[nI], fluxCellR[nI];
	#pragma unroll
	for (int q = 0; q < nI; ++q) { UL[q] = Lp[q * g.strideV]; UR[q] = Rp[q * g.strideV]; }
	Eqn::template fluxFromCons<SIDE>(fluxCellL, ep, UL);
	Eqn::template fluxFromCons<SIDE>(fluxCellR, ep, UR);
	#pragma unroll
	for (int q = 0; q < nI; ++q) {
		for (int side2 = 0; side2 < g.dim; ++side2) {
			real fL, fR;
			if (side2 == SIDE) { fL = fluxCellL[q]; fR = fluxCellR[q]; }
			else {
				real const* f = c.flux + side2 * c.blockStride + q * g.strideV + idx;
				fL = f[0];
				fR = f[ctuStride(g, side2)];
			}
			real const dF_dx = (fR * c.areaR[side2] - fL * c.areaL[side2]) * c.invVolume;
			UL[q] -= real(.5) * dt * dF_dx;
			UR[q] -= real(.5) * dt * dF_dx;
		}
		Lp[q * g.strideV] = UL[q];
		Rp[q * g.strideV] = UR[q];
	}
}

// updateCTU (ctu.cl:12-129; SETBOUNDS(1,1))
template<class Eqn, int MODE>
__global__ void ctu_update(GridP<typename Eqn::real> const g, StageP<typename Eqn::real> const sp, CtuP<typename Eqn::real> const c,
	typename Eqn::Params const ep)
{
	typedef typename Eqn::real real;
	int i, j, k; long long idx;
	if (!ctuCell(g, i, j, k, idx) || ctuOOB(g, i, j, k, 1, 1)) return;
	real const dt = real(*sp.dt);
	ctuUpdateSide<Eqn, 0>(g, c, ep, idx, dt);
	if (g.dim >= 2) ctuUpdateSide<Eqn, 1>(g, c, ep, idx, dt);
	if (g.dim >= 3) ctuUpdateSide<Eqn, 2>(g, c, ep, idx, dt);
}

// calcDerivFromFlux (fvsolver.cl:6-125) + the integrator's combination + constrainU [+ calcDT]: the epilogue of fv_stage on stored fluxes
template<class Eqn, int MODE>
__global__ void __launch_bounds__(HB_CTU_NT) ctu_finish(GridP<typename Eqn::real> const g, StageP<typename Eqn::real> const sp,
	CtuP<typename Eqn::real> const c, typename Eqn::Params const ep)
{
	typedef typename Eqn::real real;
	constexpr int nI = Eqn::nI;
	__shared__ double redBuf[HB_CTU_NT / 32];
	int i, j, k; long long idx;
	real dtCell = inf_of<real>::v();
	if (ctuCell(g, i, j, k, idx) && !ctuOOB(g, i, j, k, HB_G, HB_G)) {
		double const dt = *sp.dt;
		real acc[nI];
		#pragma unroll
		for (int q = 0; q < nI; ++q) acc[q] = 0;
		if (sp.computeL && g.volOn) {
			for (int side = 0; side < g.dim; ++side) {
				real const aov = g.aov[side];
				real const* f = c.flux + side * c.blockStride + idx;
				long long const st = ctuStride(g, side);
				#pragma unroll
				for (int q = 0; q < nI; ++q) {
					real const fl = f[q * g.strideV], fr = f[q * g.strideV + st];
					acc[q] = acc[q] - (fr * aov - fl * aov);
				}
			}
		}
		if (sp.Lout) {
			#pragma unroll
			for (int q = 0; q < nI; ++q) sp.Lout[idx + q * g.strideV] = acc[q];
		}
		if (sp.Uout) {
			real U[nI];
			#pragma unroll
			for (int q = 0; q < nI; ++q) {
				real r = 0;
				#pragma unroll
				for (int a = 0; a < HB_MAX_TERMS; ++a)
					if (a < sp.nA) r = r + sp.aPtr[a][idx + q * g.strideV] * real(sp.aCoef[a]);
				#pragma unroll
				for (int b = 0; b < HB_MAX_TERMS; ++b)
					if (b < sp.nB) r = r + sp.bPtr[b][idx + q * g.strideV] * real(sp.bCoef[b] * dt);
				if (sp.computeL) r = r + acc[q] * real(sp.betaSelf * dt);
				U[q] = r;
			}
			Eqn::constrainU(ep, U);
			#pragma unroll
			for (int q = 0; q < nI; ++q) sp.Uout[idx + q * g.strideV] = U[q];
			if (sp.dtMinBits) dtCell = rmin<real>(dtCell, Eqn::calcDTCell(ep, U, g.dx, g.dim));
		}
	}
	if (sp.dtMinBits) {
		double v = double(dtCell);
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) { double const u = __shfl_xor_sync(0xffffffffu, v, o); v = u < v ? u : v; }
		if ((threadIdx.x & 31) == 0) redBuf[threadIdx.x >> 5] = v;
		__syncthreads();
		if (threadIdx.x < 32) {
			v = threadIdx.x < HB_CTU_NT / 32 ? redBuf[threadIdx.x] : HUGE_VAL;
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) { double const u = __shfl_xor_sync(0xffffffffu, v, o); v = u < v ? u : v; }
			if (threadIdx.x == 0 && v < HUGE_VAL) atomicMin(sp.dtMinBits, dtBits(v));
		}
	}
}

enum { HB_CTUK_LR = 0, HB_CTUK_FLUX, HB_CTUK_UPDATE, HB_CTUK_FINISH };

template<class Eqn, int MODE>
cudaError_t launchCtuKernel(int which, GridP<typename Eqn::real> const& g, StageP<typename Eqn::real> const& sp, CtuP<typename Eqn::real> const& c,
	const double* eqnParams, cudaStream_t st)
{
	typename Eqn::Params const ep = Eqn::makeParams(eqnParams);
	long long const n = (long long)g.S[0] * g.S[1] * g.S[2];
	unsigned const nb = (unsigned)((n + HB_CTU_NT - 1) / HB_CTU_NT);
	switch (which) {
	case HB_CTUK_LR: ctu_calc_lr<Eqn, MODE><<<nb, HB_CTU_NT, 0, st>>>(g, sp, c); break;
	case HB_CTUK_FLUX: ctu_calc_flux<Eqn, MODE><<<nb, HB_CTU_NT, 0, st>>>(g, sp, c, ep); break;
	case HB_CTUK_UPDATE: ctu_update<Eqn, MODE><<<nb, HB_CTU_NT, 0, st>>>(g, sp, c, ep); break;
	case HB_CTUK_FINISH: ctu_finish<Eqn, MODE><<<nb, HB_CTU_NT, 0, st>>>(g, sp, c, ep); break;
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}

}   // namespace hb
