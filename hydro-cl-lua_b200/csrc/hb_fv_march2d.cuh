// hb_fv_march2d.cuh -- the fused finite-volume stage kernel for 2-D grids, one independent WARP per pencil (sm_100a).
//
// Same contract and the same per-cell arithmetic as fv_march (hb_fv_march.cuh) and fv_stage (hb_fv_kernels.cuh): one launch = one
// Runge-Kutta stage = calcLR ('plm cons', hydro/solver/plm.cl:32-91) + calcFlux (Roe, hydro/solver/fvsolver.lua:57-198,
// hydro/flux/roe.cl:17-163) + calcDerivFromFlux (hydro/solver/fvsolver.cl:6-125) + multAdd (hydro/int/rk.lua:91-112) + constrainU
// (hydro/solver/solverbase.lua:2116-2127) [+ calcDT / reduceMin, hydro/eqn/cl/calcDT.cl:38-73].  The -fmad=false build is
// bit-identical to both and to the CPU oracle (tests/test_gpu_parity.py).
//
// Why a second marching kernel.  In 2-D the only exchange between threads is along x.  fv_march routes it through shared memory
// behind one __syncthreads per row and keeps a fifth warp per CTA for the two halo columns; that warp and its registers idle ~85 % of
// the time, its SM sub-partition runs nothing else (warps 0..3 of a 4-warp CTA land on sub-partitions 0..3), and the column warps
// stall at the barrier.  Here a warp owns 32 consecutive cells of a row and marches along y on its own:
//   * lane l holds cell c = c0 - 1 + l; lanes 1..30 finish their cells, lanes 0 and 31 only supply the neighbours' face state and
//     flux (30 of 32 lanes productive instead of 4 of 5 warps);
//   * the x neighbours' face state U + .5 sigma and flux come by __shfl_up/down_sync; the y stencil lives in registers as before;
//   * every warp has its OWN 4-slot TMA ring (one cp.async.bulk.tensor per row: the 34-cell box of all nI variables) and its own
//     mbarriers, so there is no CTA-wide barrier, no shared exchange array and no idle warp: a CTA is only a container of warps;
//   * the CFL minimum leaves by one atomicMin per warp.
#pragma once
#include "hb_fv_march.cuh"
#include "hb_fv_march3.cuh"

namespace hb {

template<int NW_, int KM_, int MINB_> struct March2Cfg {
	static constexpr int NW = NW_;       // warps per CTA (independent of each other)
	static constexpr int KM = KM_;       // nominal rows per warp along the marching axis (the launcher picks the actual count, see rowsPerWarp)
	static constexpr int MINB = MINB_ & 15;
	static constexpr bool GRAV = (MINB_ & 16) != 0;   // the epilogue adds the self-gravity source (StageP::gravPot), as MarchCfg::GRAV
	static constexpr bool GEN = (MINB_ & 32) != 0;    // the general configuration (any slope limiter, no reconstruction, flux limiter, HLL / Rusanov / HLLC): March3Cfg::GEN
};

template<class C, class real> struct March2Geom {
	static constexpr int A = 16 / int(sizeof(real));           // TMA boxes start on 16-byte boundaries of the row
	static constexpr int CW = 30;                              // cells a warp finishes per row
	static constexpr int BX = 34 + (A > 2 ? A - 2 : 0);        // box: cells c0-2 .. c0+31 (+ alignment slack for float)
	static constexpr int R = 4;
	static constexpr int NT = 32 * C::NW;
	template<int nI> static HB_HD constexpr size_t slotElems() { return (size_t(nI) * BX * sizeof(real) + 127) / 128 * 128 / sizeof(real); }
	template<int nI> static HB_HD constexpr size_t warpBytes(int nOps) { return 128 + sizeof(real) * (R * slotElems<nI>() + size_t(nOps) * nI * 32); }
	template<int nI> static HB_HD constexpr size_t smemBytes(int nOps) { return (warpBytes<nI>(nOps) + 127) / 128 * 128 * C::NW + 128; }
};

template<class real> HB_D real shflUp1(real v) { return __shfl_up_sync(0xffffffffu, v, 1); }
template<class real> HB_D real shflDown1(real v) { return __shfl_down_sync(0xffffffffu, v, 1); }

template<class Eqn, int LIM, class C, int MODE>
__global__ void __launch_bounds__((March2Geom<C, typename Eqn::real>::NT), C::MINB)
fv_march2d(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ GridP<typename Eqn::real> g,
	const __grid_constant__ StageP<typename Eqn::real> sp, const __grid_constant__ typename Eqn::Params ep, int const padX, int const chunkSel, int const KM)
{
	typedef typename Eqn::real real;
	typedef March2Geom<C, real> G;
	constexpr int nI = Eqn::nI;
	constexpr int MS = 1;                                       // marching side: y
	constexpr int BX = G::BX;
	constexpr int SLOT = int(G::template slotElems<nI>());
	extern __shared__ __align__(128) unsigned char march2Smem[];
	int const tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
	size_t const wbytes = (G::template warpBytes<nI>(sp.nOps) + 127) / 128 * 128;
	unsigned char* const base = march2Smem + size_t(w) * wbytes;
	uint64_t* full = reinterpret_cast<uint64_t*>(base);         // R mbarriers of this warp
	real* ring = reinterpret_cast<real*>(base + 128);
	real* OPB = ring + G::R * SLOT;                             // staged RK operands [operand][q][lane]
	int const lim = LIM >= 0 ? LIM : sp.slopeLimiter;

	// ---- this warp's pencil: x segment and chunk of rows
	int const nSeg = (g.N[0] + G::CW - 1) / G::CW;
	// chunkSel (overlapped slab exchange, hb_fv.cu): 0 = every row, in chunks of KM; 1 = the RIM: the HB_G lowest (chunk 0) and the HB_G highest
	// (chunk 1) interior rows, whose values the neighbouring slabs need; 2 = the rows in between, in chunks of KM
	int const nm = chunkSel == 1 ? 2 : (g.N[1] - (chunkSel == 2 ? 2 * HB_G : 0) + KM - 1) / KM;
	long long const gw = (long long)blockIdx.x * C::NW + w;
	if (gw >= (long long)nSeg * nm) return;                     // whole warp leaves: nothing below synchronises across warps
	int const seg = int(gw % nSeg);
	int const bm = int(gw / nSeg);
	int kb, ke;
	if (chunkSel == 1) { kb = bm == 0 ? HB_G : g.N[1]; ke = kb + HB_G; }
	else if (chunkSel == 2) { kb = 2 * HB_G + bm * KM; ke = min(kb + KM, g.N[1]); }
	else { kb = HB_G + bm * KM; ke = min(kb + KM, HB_G + g.N[1]); }
	int const c0 = HB_G + seg * G::CW;                          // first cell this warp finishes (lane 1)
	int const gi = c0 - 1 + lane;                               // this lane's cell
	bool const inside = lane >= 1 && lane <= G::CW && gi < g.S[0] - HB_G;
	long long const colIdx = gi;
	long long const strideM = g.strideY;
	int const need = c0 - 2 + padX;                             // row element of the first cell the warp reads (cell c0 - 2)
	int const tx0 = need & ~(G::A - 1);
	int const ob = need - tx0 + 1 + lane;                       // own position inside a slot row

	uint32_t const boxBytes = uint32_t(sizeof(real) * nI * BX);
	auto issue = [&](int plane, int slot) {
		mbarExpectTx(&full[slot], boxBytes);
		tmaLoad4D(ring + slot * SLOT, &tmap, &full[slot], tx0, plane, 0, 0);
	};
	if (lane == 0) {
		for (int s = 0; s < G::R; ++s) mbarInit(&full[s], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncwarp();
	if (lane == 0) { issue(kb - 2, 0); issue(kb - 1, 1); issue(kb, 2); if (kb + 1 <= ke + 1) issue(kb + 1, 3); }

	double const dt = *sp.dt;
	real const aovX = g.aov[0], aovM = g.aov[MS];
	constexpr bool GEN = C::GEN;
	// GEN: 0 no reconstruction, 1 'plm cons', 2 Roe with a flux limiter on cell-centred states (as fv_march3)
	int const gmode = !GEN ? 1 : (sp.plmMode == 1 ? 1 : (sp.fluxLimiter > 0 ? 2 : 0));
	real const dtR = GEN ? real(dt) : real(0);
	real Um[nI], zfP[nI], FzP[nI], accP[nI];
	real Umm[GEN ? nI : 1];                              // GEN, flux limiter: U[k-2] of the own column
	mbarWait(&full[0], 0);
	#pragma unroll
	for (int q = 0; q < nI; ++q) { Um[q] = ring[q * BX + ob]; zfP[q] = 0; FzP[q] = 0; accP[q] = 0; if constexpr (GEN) Umm[q] = Um[q]; }
	mbarWait(&full[1], 0);

	real dtCell = inf_of<real>::v(), rateCell = 0;
	int const OPS = 32;
	for (int k = kb - 1, it = 0; k <= ke; ++k, ++it) {
		int const sP = it & 3, sK = (it + 1) & 3, sN = (it + 2) & 3;
		uint32_t const parN = uint32_t((it + 2) >> 2) & 1u;
		bool const xy = k >= kb && k < ke;
		real const* __restrict__ P = ring + sK * SLOT;
		mbarWait(&full[sN], parN);
		// plane k-1 is not read any more (it lives in Um): its slot takes plane k+3
		__syncwarp();
		if (lane == 0 && k + 3 <= ke + 1) {
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			issue(k + 3, sP);
		}
		long long const idxK = colIdx + strideM * k;
		real Uk[nI], Fz[nI], UR[nI], zfN[nI];
		if constexpr (GEN) {
			real Un[nI];
			#pragma unroll
			for (int q = 0; q < nI; ++q) {
				Uk[q] = P[q * BX + ob];
				Un[q] = ring[sN * SLOT + q * BX + ob];
				real const sK = gmode == 1 ? plmHalfSlope<real>(lim, Um[q], Uk[q], Un[q]) : real(0);
				UR[q] = gmode == 1 ? Uk[q] - sK : Uk[q];
				zfN[q] = gmode == 1 ? Uk[q] + sK : Uk[q];
				Fz[q] = 0;
			}
			if (k >= kb && g.fluxOn[MS]) {
				if (gmode == 2) roeFluxLimited<Eqn, MS>(Fz, ep, sp.fluxLimiter, dtR / g.dx[MS], Umm, Um, Uk, Un);
				else interfaceFlux<Eqn, MS>(sp.flux, sp.fluxParam, Fz, ep, gmode == 1 ? zfP : Um, UR);
			}
		} else {
		#pragma unroll
		for (int q = 0; q < nI; ++q) {
			Uk[q] = P[q * BX + ob];
			plmCellFacesT<real, LIM, Eqn::FAST>(lim, Um[q], Uk[q], ring[sN * SLOT + q * BX + ob], UR[q], zfN[q]);
			Fz[q] = 0;
		}
		if (k >= kb && g.fluxOn[MS]) roeFluxAuto<Eqn, MS>(Fz, ep, zfP, UR);
		}
		if (k > kb && inside) {
			real acc[nI];
			#pragma unroll
			for (int q = 0; q < nI; ++q) acc[q] = g.volOn ? accP[q] - (Fz[q] * aovM - FzP[q] * aovM) : real(0);
			cpAsyncWaitAll();
			stageEpilogue3<Eqn, C::GRAV>(g, sp, ep, idxK - strideM, acc, Um, dt, dtCell, rateCell, OPB + lane, OPS);
		}
		if (inside && xy && sp.Uout) {
			#pragma unroll 1
			for (int o = 0; o < sp.nOps; ++o) {
				real const* src = sp.opPtr[o] + idxK;
				real* dst = OPB + o * (nI * OPS) + lane;
				#pragma unroll
				for (int q = 0; q < nI; ++q) cpAsyncElem<real>(dst + q * OPS, src + q * g.strideV);
			}
			cpAsyncCommit();
		}
		// ---- x: half slope of the own cell, the left neighbour's right face by shuffle, Roe flux at the own low face, the right
		// neighbour's flux by shuffle, flux difference (fvsolver.cl:97-123)
		#pragma unroll
		for (int q = 0; q < nI; ++q) accP[q] = 0;
		if (GEN && xy) {
			// the low-face flux straight from the row's four-cell stencil (lane 0 has no cell c0 - 3 in its box: it takes lane 1's stencil, its
			// flux is never used)
			real F[nI];
			if (g.fluxOn[0]) lowFaceFlux<Eqn, 0, LIM, BX, true>(F, ep, lim, P, lane == 0 ? ob + 1 : ob, 1, gmode, sp.flux, sp.fluxParam, sp.fluxLimiter, dtR / g.dx[0]);
			else {
				#pragma unroll
				for (int q = 0; q < nI; ++q) F[q] = 0;
			}
			#pragma unroll
			for (int q = 0; q < nI; ++q) {
				real const Fhi = shflDown1<real>(F[q]);
				if (g.volOn) accP[q] = real(0) - (Fhi * aovX - F[q] * aovX);
			}
		} else if (xy) {                                        // warp-uniform
			real UL[nI], URx[nI], F[nI];
			#pragma unroll
			for (int q = 0; q < nI; ++q) {
				real faceR;
				plmCellFacesT<real, LIM, Eqn::FAST>(lim, P[q * BX + ob - 1], Uk[q], P[q * BX + ob + 1], URx[q], faceR);
				UL[q] = shflUp1<real>(faceR);
				if (lane == 0) UL[q] = URx[q];                  // lane 0 has no left neighbour in the warp: its flux is never used
			}
			if (g.fluxOn[0]) roeFluxAuto<Eqn, 0>(F, ep, UL, URx);
			else {
				#pragma unroll
				for (int q = 0; q < nI; ++q) F[q] = 0;
			}
			#pragma unroll
			for (int q = 0; q < nI; ++q) {
				real const Fhi = shflDown1<real>(F[q]);
				if (g.volOn) accP[q] = real(0) - (Fhi * aovX - F[q] * aovX);
			}
		}
		#pragma unroll
		for (int q = 0; q < nI; ++q) { if constexpr (GEN) Umm[q] = Um[q]; Um[q] = Uk[q]; zfP[q] = zfN[q]; FzP[q] = Fz[q]; }
	}
	if (sp.dtMinBits) {
		if (rateCell > real(0)) dtCell = rmin<real>(dtCell, real(1.) / rateCell);
		double v = double(dtCell);
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) { double const u = __shfl_xor_sync(0xffffffffu, v, o); v = u < v ? u : v; }
		if (lane == 0 && v < HUGE_VAL) atomicMin(sp.dtMinBits, dtBits(v));
	}
}

}   // namespace hb
