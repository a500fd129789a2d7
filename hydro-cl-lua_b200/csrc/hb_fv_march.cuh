// hb_fv_march.cuh -- the plane-marching fused finite-volume stage kernel (sm_100a; TMA + mbarrier ring).
//
// Same contract as fv_stage (hb_fv_kernels.cuh): one launch = one Runge-Kutta stage =
//     calcLR ('plm cons', hydro/solver/plm.cl:32-91) + calcFlux (Roe, hydro/solver/fvsolver.lua:57-198, hydro/flux/roe.cl:17-163)
//     + calcDerivFromFlux (hydro/solver/fvsolver.cl:6-125) + the stage's multAdd combination (hydro/int/rk.lua:91-112)
//     + constrainU (hydro/solver/solverbase.lua:2116-2127) [+ calcDT / reduceMin on the last stage, hydro/eqn/cl/calcDT.cl:38-73]
// but organised for the B200's real bottleneck on this path, the FP64 pipe and the instruction issue slots, not HBM:
//
//   * 2.5-D blocking.  A CTA owns a TX x TY column tile (x on lanes, y on warps) and marches along the slowest axis
//     (z in 3-D, y in 2-D).  One thread owns one column: the marching-axis stencil (U[k-1], U[k], U[k+1], the face state
//     and the flux of the previous interface) lives in registers, so that axis costs no shared-memory traffic and no
//     redundant halo work except one extra interface flux per KM planes.
//   * TMA.  Plane k+3 of the tile plus its 2-cell x/y halo (all nI variables, one 4-D box) is fetched by one
//     cp.async.bulk.tensor instruction into a 4-slot shared-memory ring while plane k is being computed; completion is
//     signalled on an mbarrier per slot.  Out-of-range box parts are zero-filled by the hardware (those lanes never store).
//   * Whole-warp halos.  The x/y neighbours' half slopes and interface fluxes are exchanged through shared memory; the halo
//     work that no column thread owns (slopes of the cells one step outside the tile, the flux of the tile's far faces) is
//     done by extra warps (two row warps, one x-halo warp), so no lane of a working warp idles on tile overlap.
//   * The slope limiter is a template parameter: the 20-way switch of hydro/app.lua:614-635 (with its divisions) is not
//     replicated 30 times in the instruction stream, which keeps the kernel inside the instruction cache.
//
// Arithmetic per cell is the same sequence of operations as fv_stage (and as the CPU oracle): slopes, faces U -/+ .5 sigma,
// Roe flux, acc = ((0 - dFx) - dFy) - dFz, RK combination alpha terms then beta terms, constrainU.  The -fmad=false build of
// this kernel is therefore bit-identical to the oracle as well (tests/test_gpu_parity.py).
#pragma once
#include "hb_rtc_compat.h"
#include "hb_fv_kernels.cuh"
#include "hb_roe_fast.cuh"

namespace hb {

template<int WX_, int TY_, int KM_, int MINB_, int VAR_ = 0> struct MarchCfg {
	static constexpr int WX = WX_;     // warps along x: TX = 32 * WX
	static constexpr int TY = TY_;     // rows (3-D only)
	static constexpr int KM = KM_;     // planes per CTA along the marching axis
	static constexpr int MINB = MINB_; // __launch_bounds__ min blocks per SM
	// instruction-level parallelism of the column warps: 0 = the three interface fluxes of a cell one after the other;
	// 1 = all of them issued as one straight-line block (the cores interleave: three independent dependent chains);
	// 2 = marching-axis flux alone, then the x and y fluxes as a pair
	static constexpr int VAR = VAR_ & 15;
	static constexpr bool STAGGER = (VAR_ & 16) != 0;   // column warps w, w + 4 run the slope phase at opposite ends of the iteration
	static constexpr bool GRAV = (VAR_ & 32) != 0;      // the epilogue adds the self-gravity source (StageP::gravPot) to L: a separate configuration, so that
	                                                    // the default kernel's register allocation is untouched
};

template<int DIM, class C, class real> struct MarchGeom {
	static constexpr int WX = C::WX;
	static constexpr int TX = 32 * C::WX;
	static constexpr int TY = DIM == 3 ? C::TY : 1;
	static constexpr int GYB = DIM == 3 ? HB_G : 0;
	// x halo of the TMA box: at least the 2 ghost cells, widened so that the box starts on a 16-byte boundary of the row
	// (the hardware rejects a tile whose innermost start is not 16-byte aligned: float needs 4 cells)
	static constexpr int HL = (16 / int(sizeof(real))) > HB_G ? (16 / int(sizeof(real))) : HB_G;
	static constexpr int BX = TX + 2 * HL, BY = TY + 2 * GYB;
	static constexpr int PS = BX * BY;                         // variable stride inside a ring slot
	static constexpr int NREG = TY * WX;                       // column warps
	static constexpr int NROWH = DIM == 3 ? 2 * WX : 0;        // row-halo warps (row -1, row TY)
	static constexpr int NWARPS = NREG + NROWH + 1;            // + the x-halo warp
	static constexpr int NT = 32 * NWARPS;
	static constexpr int R = 4;                                // ring slots: planes k-1 (being recycled), k, k+1, k+2 (in flight)
	static constexpr int SGXN = TY * (TX + 2), FXXN = TY * (TX + 1);
	static constexpr int SGYN = DIM == 3 ? (TY + 2) * TX : 0, FXYN = DIM == 3 ? (TY + 1) * TX : 0;
	template<int nI> static constexpr size_t slotBytes() { return (sizeof(real) * nI * PS + 127) / 128 * 128; }
	// nOps = number of RK operands (alpha terms other than the stage input + beta terms) staged per column thread
	template<int nI> static constexpr size_t smemBytes(int nOps) {
		return 128 + R * slotBytes<nI>() + sizeof(real) * nI * size_t(2 * (SGXN + FXXN + SGYN + FXYN) + nOps * NREG * 32) + 128;
	}
	static_assert(TY <= 16, "the x-halo warp serves at most 16 rows");
};

// ---- PTX wrappers (mbarrier + TMA tile load)
HB_D uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
HB_D void mbarInit(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smemAddr(bar)), "r"(count) : "memory");
}
HB_D void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smemAddr(bar)), "r"(bytes) : "memory");
}
HB_D void mbarWait(uint64_t* bar, uint32_t parity) {
	asm volatile(
		"{\n"
		".reg .pred P1;\n"
		"LAB_WAIT:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
		"@P1 bra DONE;\n"
		"bra LAB_WAIT;\n"
		"DONE:\n"
		"}" :: "r"(smemAddr(bar)), "r"(parity) : "memory");
}
// per-thread asynchronous global -> shared copy of one element (LDGSTS): no register is held while the load is in flight
template<class real> HB_D void cpAsyncElem(real* dstSmem, const real* src) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" :: "r"(smemAddr(dstSmem)), "l"(src), "n"(int(sizeof(real))) : "memory");
}
HB_D void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
HB_D void cpAsyncWaitAll() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
HB_D void tmaLoad4D(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
	asm volatile(
		"cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
		:: "r"(smemAddr(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smemAddr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
		: "memory");
}

// RK combination + constrainU + stores + CFL dt of one finished cell (hydro/int/rk.lua:96-112, solverbase.lua:2116-2127).
// `own` = the stage input state of this cell (register copy), used for alpha terms that point at the stage input.
// `ops` = this thread's staged RK operands in shared memory, [operand][q] with stride opStride between entries (see fv_march).
template<class Eqn, bool GRAV = false>
HB_D void stageEpilogue(GridP<typename Eqn::real> const& g, StageP<typename Eqn::real> const& sp, typename Eqn::Params const& ep,
	long long idx, typename Eqn::real const (&accIn)[Eqn::nI], typename Eqn::real const (&own)[Eqn::nI], double dt,
	typename Eqn::real& dtCell, typename Eqn::real& rateCell, typename Eqn::real const* ops, int opStride)
{
	typedef typename Eqn::real real;
	constexpr int nI = Eqn::nI;
	real accG[nI];
	real const* acc = accIn;
	if constexpr (GRAV && Eqn::eqnId <= 1) {
		// op:addSource of the self-gravity op (solverbase.lua:3219-3223, selfgrav.cl:11-76), as in fv_stage: accel = central difference of the
		// potential, deriv.m -= accel rho, deriv.ETotal -= m . accel
		#pragma unroll
		for (int q = 0; q < nI; ++q) accG[q] = accIn[q];
		if (sp.gravPot && sp.computeL) {
			real accel[3] = {0, 0, 0};
			accel[0] = (sp.gravPot[idx + 1] - sp.gravPot[idx - 1]) / (real(2.) * g.dx[0]);
			if (g.dim >= 2) accel[1] = (sp.gravPot[idx + g.strideY] - sp.gravPot[idx - g.strideY]) / (real(2.) * g.dx[1]);
			if (g.dim >= 3) accel[2] = (sp.gravPot[idx + g.strideZ] - sp.gravPot[idx - g.strideZ]) / (real(2.) * g.dx[2]);
			accG[1] = accG[1] - accel[0] * own[0];
			accG[2] = accG[2] - accel[1] * own[0];
			accG[3] = accG[3] - accel[2] * own[0];
			accG[4] -= own[1] * accel[0] + own[2] * accel[1] + own[3] * accel[2];
		}
		acc = accG;
	}
	if (sp.Lout) {
		#pragma unroll
		for (int q = 0; q < nI; ++q) sp.Lout[idx + q * g.strideV] = acc[q];
	}
	if (!sp.Uout) return;
	real U[nI];
	#pragma unroll
	for (int q = 0; q < nI; ++q) U[q] = 0;
	int slot = 0;
	#pragma unroll
	for (int a = 0; a < HB_MAX_TERMS; ++a)
		if (a < sp.nA) {
			real const c = real(sp.aCoef[a]);
			if ((sp.aOwnMask >> a) & 1) {
				#pragma unroll
				for (int q = 0; q < nI; ++q) U[q] = U[q] + own[q] * c;
			} else {
				#pragma unroll
				for (int q = 0; q < nI; ++q) U[q] = U[q] + ops[(slot * nI + q) * opStride] * c;
				++slot;
			}
		}
	#pragma unroll
	for (int b = 0; b < HB_MAX_TERMS; ++b)
		if (b < sp.nB) {
			real const c = real(sp.bCoef[b] * dt);
			#pragma unroll
			for (int q = 0; q < nI; ++q) U[q] = U[q] + ops[(slot * nI + q) * opStride] * c;
			++slot;
		}
	if (sp.computeL) {
		real const c = real(sp.betaSelf * dt);
		#pragma unroll
		for (int q = 0; q < nI; ++q) U[q] = U[q] + acc[q] * c;
	}
	finishCellAuto<Eqn>(ep, U, g.dx, g.invdx, g.dim, sp.dtMinBits != nullptr, dtCell, rateCell);
	#pragma unroll
	for (int q = 0; q < nI; ++q) sp.Uout[idx + q * g.strideV] = U[q];
}

template<class Eqn, int DIM, int LIM, class C, int MODE>
__global__ void __launch_bounds__((MarchGeom<DIM, C, typename Eqn::real>::NT), C::MINB)
fv_march(const __grid_constant__ CUtensorMap tmap, GridP<typename Eqn::real> const g, StageP<typename Eqn::real> const sp,
	typename Eqn::Params const ep, int const padX, int const chunkSel)
{
	typedef typename Eqn::real real;
	typedef MarchGeom<DIM, C, real> G;
	constexpr int nI = Eqn::nI;
	constexpr int MS = DIM - 1;                 // marching side
	constexpr int TX = G::TX, TY = G::TY, BX = G::BX, PS = G::PS;
	constexpr int SLOT = (int(sizeof(real)) * nI * PS + 127) / 128 * 128 / int(sizeof(real));   // == slotBytes / sizeof(real)
	extern __shared__ __align__(128) unsigned char marchSmem[];
	uint64_t* full = reinterpret_cast<uint64_t*>(marchSmem);               // R mbarriers
	real* ring = reinterpret_cast<real*>(marchSmem + 128);
	// exchange arrays, two halves each (plane parity): one barrier per iteration suffices (see the loop)
	real* SGX = ring + G::R * SLOT;             // half slopes along x of cells i = -1 .. TX      [parity][q][row][i + 1]
	real* FXX = SGX + 2 * nI * G::SGXN;         // x fluxes at the low faces of cells i = 0 .. TX [parity][q][row][i]
	real* SGY = FXX + 2 * nI * G::FXXN;         // half slopes along y of rows j = -1 .. TY       [parity][q][j + 1][i]
	real* FXY = SGY + 2 * nI * G::SGYN;         // y fluxes at the low faces of rows j = 0 .. TY  [parity][q][j][i]
	real* OPB = FXY + 2 * nI * G::FXYN;         // staged RK operands of the column threads       [operand][q][thread]
	constexpr int OPS = G::NREG * 32;
	__shared__ double redBuf[32];

	int const tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
	int const lim = LIM >= 0 ? LIM : sp.slopeLimiter;

	// ---- tile
	int const ntx = (g.N[0] + TX - 1) / TX;
	int const nty = DIM == 3 ? (g.N[1] + TY - 1) / TY : 1;
	int bid = blockIdx.x;
	int const bx = bid % ntx; bid /= ntx;
	int const by = bid % nty; int bm = bid / nty;
	// chunkSel (overlapped slab exchange, hb_fv.cu): 0 = every chunk of KM planes; 1 = the first and the last chunk (whose planes are
	// sent to the neighbours); 2 = the chunks in between
	if (chunkSel == 1) { if (bm != 0) bm = (g.N[MS] + C::KM - 1) / C::KM - 1; }
	else if (chunkSel == 2) bm += 1;
	int const i0 = bx * TX + HB_G;
	int const j0 = DIM == 3 ? by * TY + HB_G : 0;
	int const kb = bm * C::KM + HB_G;
	int const ke = min(kb + C::KM, HB_G + g.N[MS]);      // exclusive

	// ---- role of this thread
	int ci, cj;
	bool doMain = false, doSX = false, doFX = false, doSY = false, doFY = false;
	if (w < G::NREG) {
		cj = w / G::WX; ci = (w % G::WX) * 32 + lane;
		doMain = doSX = doFX = true; doSY = doFY = DIM == 3;
	} else if (DIM == 3 && w < G::NREG + G::WX) {
		cj = -1; ci = (w - G::NREG) * 32 + lane; doSY = true;
	} else if (DIM == 3 && w < G::NREG + 2 * G::WX) {
		cj = TY; ci = (w - G::NREG - G::WX) * 32 + lane; doSY = doFY = true;
	} else {
		int const r = lane & 15; bool const right = lane >= 16;
		if (r < TY) { cj = r; ci = right ? TX : -1; doSX = true; doFX = right; }
		else { cj = 0; ci = 0; }
	}
	int const ob = (cj + G::GYB) * BX + (ci + G::HL);      // own position inside a slot (per variable)
	int const gi = i0 + ci, gj = j0 + cj;
	bool const inside = doMain && gi < g.S[0] - HB_G && (DIM < 3 || gj < g.S[1] - HB_G);
	long long const colIdx = DIM == 3 ? gi + g.strideY * gj : gi;
	long long const strideM = DIM == 3 ? g.strideZ : g.strideY;

	// ---- TMA ring
	uint32_t const boxBytes = uint32_t(sizeof(real) * nI * PS);
	int const tx0 = i0 - G::HL + padX, ty0 = DIM == 3 ? j0 - HB_G : 0;
	auto issue = [&](int plane, int slot) {
		mbarExpectTx(&full[slot], boxBytes);
		if (DIM == 3) tmaLoad4D(ring + slot * SLOT, &tmap, &full[slot], tx0, ty0, plane, 0);
		else tmaLoad4D(ring + slot * SLOT, &tmap, &full[slot], tx0, plane, 0, 0);
	};
	if (tid == 0) {
		for (int s = 0; s < G::R; ++s) mbarInit(&full[s], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads();
	if (tid == 0) { issue(kb - 2, 0); issue(kb - 1, 1); issue(kb, 2); if (kb + 1 <= ke + 1) issue(kb + 1, 3); }

	double const dt = *sp.dt;
	real const aovX = g.aov[0], aovY = g.aov[1], aovM = g.aov[MS];
	real Um[nI], zfP[nI], FzP[nI], accP[nI];          // U[k-1], face state of cell k-1 towards k, flux at k-3/2, x/y part of dU/dt[k-1]
	mbarWait(&full[0], 0);
	#pragma unroll
	for (int q = 0; q < nI; ++q) { Um[q] = doMain ? ring[q * PS + ob] : real(0); zfP[q] = 0; FzP[q] = 0; accP[q] = 0; }
	mbarWait(&full[1], 0);

	real dtCell = inf_of<real>::v(), rateCell = 0;
	bool const earlySlopes = C::STAGGER && w < G::NREG && ((w >> 2) & 1);
	// iteration `it` handles plane k = kb - 1 + it.  Plane p occupies ring slot (p - (kb-2)) % 4 on its ((p - (kb-2)) / 4)-th use.
	// Plane k+3 is requested at the barrier of iteration k and first needed at the top of iteration k+2: a full iteration of lead.
	for (int k = kb - 1, it = 0; k <= ke; ++k, ++it) {
		int const sP = it & 3, sK = (it + 1) & 3, sN = (it + 2) & 3;     // slots of planes k-1 (recycled for k+3), k, k+1
		uint32_t const parN = uint32_t((it + 2) >> 2) & 1u;
		bool const xy = k >= kb && k < ke;
		real const* __restrict__ P = ring + sK * SLOT;
		mbarWait(&full[sN], parN);
		real Uk[nI];                                   // own cell of plane k (every role reads it at least once)
		#pragma unroll
		for (int q = 0; q < nI; ++q) Uk[q] = P[q * PS + ob];
		// ---- half slopes of plane k+1 along x and y (plm.cl:56-76), published for the next iteration (other buffer half).
		// Independent of the fluxes of plane k: the column warps that share an SM sub-partition with another column warp (w and w + 4)
		// run it at opposite ends of the iteration, so that one warp's FP64-dense flux phase overlaps the other's shared-memory-bound
		// phases instead of both competing for the FP64 pipe and then both leaving it idle.
		auto slopesPhase = [&]() {
			if (k + 1 >= kb && k + 1 < ke) {
				real const* __restrict__ Q = ring + sN * SLOT;
				real* const sgxN = SGX + ((k + 1) & 1) * (nI * G::SGXN);
				real* const sgyN = SGY + ((k + 1) & 1) * (nI * G::SGYN);
				// all loads first, then all stores: the compiler cannot move a shared-memory load above a shared-memory store, and
				// a load -> slope -> store sequence per variable would pay the shared-memory latency nI times in a row
				if (DIM == 3 && doSX && doSY) {
					real sx[nI], sy[nI];
					#pragma unroll
					for (int q = 0; q < nI; ++q) {
						real const c = Q[q * PS + ob];
						sx[q] = plmHalfSlopeT<real, LIM, Eqn::FAST>(lim, Q[q * PS + ob - 1], c, Q[q * PS + ob + 1]);
						sy[q] = plmHalfSlopeT<real, LIM, Eqn::FAST>(lim, Q[q * PS + ob - BX], c, Q[q * PS + ob + BX]);
					}
					#pragma unroll
					for (int q = 0; q < nI; ++q) {
						sgxN[(q * TY + cj) * (TX + 2) + ci + 1] = sx[q];
						sgyN[(q * (TY + 2) + cj + 1) * TX + ci] = sy[q];
					}
				} else if (doSX) {
					real sx[nI];
					#pragma unroll
					for (int q = 0; q < nI; ++q) sx[q] = plmHalfSlopeT<real, LIM, Eqn::FAST>(lim, Q[q * PS + ob - 1], Q[q * PS + ob], Q[q * PS + ob + 1]);
					#pragma unroll
					for (int q = 0; q < nI; ++q) sgxN[(q * TY + cj) * (TX + 2) + ci + 1] = sx[q];
				} else if (DIM == 3 && doSY) {
					real sy[nI];
					#pragma unroll
					for (int q = 0; q < nI; ++q) sy[q] = plmHalfSlopeT<real, LIM, Eqn::FAST>(lim, Q[q * PS + ob - BX], Q[q * PS + ob], Q[q * PS + ob + BX]);
					#pragma unroll
					for (int q = 0; q < nI; ++q) sgyN[(q * (TY + 2) + cj + 1) * TX + ci] = sy[q];
				}
			}
		};
		if (earlySlopes) slopesPhase();
		// ---- column warps, fused form (MarchCfg::VAR >= 1): the fluxes at the low z, x and y faces of cell (ci, cj, k) are independent
		// of each other, so their cores are issued as one straight-line block and the scheduler interleaves the dependent chains
		// (MUFU seed -> Newton steps -> wave strengths).  Same operations per flux as the separate form below.
		bool const fusedMain = C::VAR >= 1 && w < G::NREG;
		if (C::VAR >= 1 && fusedMain) {
			long long const idxK = colIdx + strideM * k;
			real Fz[nI], UR[nI], zfN[nI];
			#pragma unroll
			for (int q = 0; q < nI; ++q) {
				real const s = plmHalfSlopeT<real, LIM, Eqn::FAST>(lim, Um[q], Uk[q], ring[sN * SLOT + q * PS + ob]);
				UR[q] = Uk[q] - s;
				zfN[q] = Uk[q] + s;
				Fz[q] = 0;
			}
			if (xy) {
				real* const fxx = FXX + (k & 1) * (nI * G::FXXN);
				real* const fxy = FXY + (k & 1) * (nI * G::FXYN);
				real const* const sgx = SGX + (k & 1) * (nI * G::SGXN);
				real const* const sgy = SGY + (k & 1) * (nI * G::SGYN);
				real ULx[nI], URx[nI], ULy[nI], URy[nI], Fx[nI], Fy[nI];
				#pragma unroll
				for (int q = 0; q < nI; ++q) {
					real const* sg = sgx + (q * TY + cj) * (TX + 2) + ci;
					ULx[q] = P[q * PS + ob - 1] + sg[0];
					URx[q] = P[q * PS + ob] - sg[1];
					if (DIM == 3) {
						real const* sh = sgy + (q * (TY + 2) + cj) * TX + ci;
						ULy[q] = P[q * PS + ob - BX] + sh[0];
						URy[q] = P[q * PS + ob] - sh[TX];
					}
				}
				if constexpr (DIM == 3 && C::VAR == 3 && Eqn::FAST && Eqn::eqnId == 0) {
					// z flux on its own; x and y cores as one block; a state pair on one of the reference's special branches (rare) is
					// redone by the literal code from operands re-read from shared memory, so no operand stays live across the cores
					roeFluxAuto<Eqn, MS, DIM == 3>(Fz, ep, zfP, UR);
					bool const rx = eulerRoeFluxCore<Eqn, 0>(Fx, ep, ULx, URx);
					bool const ry = eulerRoeFluxCore<Eqn, 1>(Fy, ep, ULy, URy);
					if (!(rx && ry)) {
						real A[nI], B[nI];
						#pragma unroll
						for (int q = 0; q < nI; ++q) {
							real const* sg = sgx + (q * TY + cj) * (TX + 2) + ci;
							A[q] = P[q * PS + ob - 1] + sg[0];
							B[q] = P[q * PS + ob] - sg[1];
						}
						eulerRoeFluxFixup<Eqn, 0>(rx, Fx, ep, A, B);
						#pragma unroll
						for (int q = 0; q < nI; ++q) {
							real const* sh = sgy + (q * (TY + 2) + cj) * TX + ci;
							A[q] = P[q * PS + ob - BX] + sh[0];
							B[q] = P[q * PS + ob] - sh[TX];
						}
						eulerRoeFluxFixup<Eqn, 1>(ry, Fy, ep, A, B);
					}
				} else if constexpr (DIM == 3) {
					if constexpr (C::VAR == 1) roeFluxTripleAuto<Eqn, MS, 0, 1>(Fz, Fx, Fy, ep, zfP, UR, ULx, URx, ULy, URy);
					else {
						roeFluxAuto<Eqn, MS, DIM == 3>(Fz, ep, zfP, UR);
						roeFluxPairAuto<Eqn, 0, 1>(Fx, Fy, ep, ULx, URx, ULy, URy);
					}
				} else {
					roeFluxPairAuto<Eqn, MS, 0>(Fz, Fx, ep, zfP, UR, ULx, URx);
				}
				bool const onM = g.fluxOn[MS], on0 = g.fluxOn[0], on1 = g.fluxOn[1];
				#pragma unroll
				for (int q = 0; q < nI; ++q) {
					if (!onM) Fz[q] = 0;
					fxx[(q * TY + cj) * (TX + 1) + ci] = on0 ? Fx[q] : real(0);
					if (DIM == 3) fxy[(q * (TY + 1) + cj) * TX + ci] = on1 ? Fy[q] : real(0);
				}
			} else if (k >= kb && g.fluxOn[MS]) roeFluxAuto<Eqn, MS, DIM == 3>(Fz, ep, zfP, UR);
			if (k > kb && inside) {
				real acc[nI];
				#pragma unroll
				for (int q = 0; q < nI; ++q) acc[q] = g.volOn ? accP[q] - (Fz[q] * aovM - FzP[q] * aovM) : real(0);
				cpAsyncWaitAll();
				stageEpilogue<Eqn, C::GRAV>(g, sp, ep, idxK - strideM, acc, Um, dt, dtCell, rateCell, OPB + tid, OPS);
			}
			if (inside && xy && sp.Uout) {
				int slot = 0;
				#pragma unroll
				for (int a = 0; a < HB_MAX_TERMS; ++a)
					if (a < sp.nA && !((sp.aOwnMask >> a) & 1)) {
						#pragma unroll
						for (int q = 0; q < nI; ++q) cpAsyncElem<real>(OPB + (slot * nI + q) * OPS + tid, sp.aPtr[a] + idxK + q * g.strideV);
						++slot;
					}
				#pragma unroll
				for (int b = 0; b < HB_MAX_TERMS; ++b)
					if (b < sp.nB) {
						#pragma unroll
						for (int q = 0; q < nI; ++q) cpAsyncElem<real>(OPB + (slot * nI + q) * OPS + tid, sp.bPtr[b] + idxK + q * g.strideV);
						++slot;
					}
				cpAsyncCommit();
			}
			#pragma unroll
			for (int q = 0; q < nI; ++q) { Um[q] = Uk[q]; zfP[q] = zfN[q]; FzP[q] = Fz[q]; }
		}
		// ---- marching axis, registers only: slope of cell k, flux at interface k-1/2, finish cell k-1
		if (doMain && !fusedMain) {
			long long const idxK = colIdx + strideM * k;
			real Fz[nI], UR[nI], zfN[nI];
			#pragma unroll
			for (int q = 0; q < nI; ++q) {
				real const s = plmHalfSlopeT<real, LIM, Eqn::FAST>(lim, Um[q], Uk[q], ring[sN * SLOT + q * PS + ob]);
				UR[q] = Uk[q] - s;
				zfN[q] = Uk[q] + s;
				Fz[q] = 0;
			}
			if (k >= kb && g.fluxOn[MS]) roeFluxAuto<Eqn, MS, DIM == 3>(Fz, ep, zfP, UR);
			if (k > kb && inside) {
				real acc[nI];
				#pragma unroll
				for (int q = 0; q < nI; ++q) acc[q] = g.volOn ? accP[q] - (Fz[q] * aovM - FzP[q] * aovM) : real(0);
				cpAsyncWaitAll();
				stageEpilogue<Eqn, C::GRAV>(g, sp, ep, idxK - strideM, acc, Um, dt, dtCell, rateCell, OPB + tid, OPS);
			}
			if (inside && xy && sp.Uout) {
				// the RK operands of cell k (alpha terms other than the stage input, beta terms) are consumed one iteration from
				// now: start their global -> shared copies (per-thread slots, so no barrier is involved)
				int slot = 0;
				#pragma unroll
				for (int a = 0; a < HB_MAX_TERMS; ++a)
					if (a < sp.nA && !((sp.aOwnMask >> a) & 1)) {
						#pragma unroll
						for (int q = 0; q < nI; ++q) cpAsyncElem<real>(OPB + (slot * nI + q) * OPS + tid, sp.aPtr[a] + idxK + q * g.strideV);
						++slot;
					}
				#pragma unroll
				for (int b = 0; b < HB_MAX_TERMS; ++b)
					if (b < sp.nB) {
						#pragma unroll
						for (int q = 0; q < nI; ++q) cpAsyncElem<real>(OPB + (slot * nI + q) * OPS + tid, sp.bPtr[b] + idxK + q * g.strideV);
						++slot;
					}
				cpAsyncCommit();
			}
			#pragma unroll
			for (int q = 0; q < nI; ++q) { Um[q] = Uk[q]; zfP[q] = zfN[q]; FzP[q] = Fz[q]; }
		}
		// ---- fluxes of plane k: Roe flux at the low x and y faces (slopes of plane k were published one iteration ago)
		real* const sgx = SGX + (k & 1) * (nI * G::SGXN);
		real* const sgy = SGY + (k & 1) * (nI * G::SGYN);
		real* const fxx = FXX + (k & 1) * (nI * G::FXXN);
		real* const fxy = FXY + (k & 1) * (nI * G::FXYN);
		if (xy && !fusedMain) {
			if (doFX) {
				real F[nI];
				if (g.fluxOn[0]) {
					real UL[nI], UR[nI];
					#pragma unroll
					for (int q = 0; q < nI; ++q) {
						real const* sg = sgx + (q * TY + cj) * (TX + 2) + ci;
						UL[q] = P[q * PS + ob - 1] + sg[0];
						UR[q] = P[q * PS + ob] - sg[1];
					}
					roeFluxAuto<Eqn, 0, DIM == 3>(F, ep, UL, UR);
				} else {
					#pragma unroll
					for (int q = 0; q < nI; ++q) F[q] = 0;
				}
				#pragma unroll
				for (int q = 0; q < nI; ++q) fxx[(q * TY + cj) * (TX + 1) + ci] = F[q];
			}
			if (DIM == 3 && doFY) {
				real F[nI];
				if (g.fluxOn[1]) {
					real UL[nI], UR[nI];
					#pragma unroll
					for (int q = 0; q < nI; ++q) {
						real const* sg = sgy + (q * (TY + 2) + cj) * TX + ci;
						UL[q] = P[q * PS + ob - BX] + sg[0];
						UR[q] = P[q * PS + ob] - sg[TX];
					}
					roeFluxAuto<Eqn, 1, DIM == 3>(F, ep, UL, UR);
				} else {
					#pragma unroll
					for (int q = 0; q < nI; ++q) F[q] = 0;
				}
				#pragma unroll
				for (int q = 0; q < nI; ++q) fxy[(q * (TY + 1) + cj) * TX + ci] = F[q];
			}
		}
		if (!earlySlopes) slopesPhase();
		__syncthreads();           // the only barrier of the iteration: fluxes of plane k and slopes of plane k+1 are visible
		if (tid == 0 && k + 3 <= ke + 1) {
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			issue(k + 3, sP);
		}
		// ---- flux differences of plane k along x and y (fvsolver.cl:97-123)
		if (doMain) {
			#pragma unroll
			for (int q = 0; q < nI; ++q) accP[q] = 0;
			if (xy && g.volOn) {
				#pragma unroll
				for (int q = 0; q < nI; ++q) {
					real const* fx = fxx + (q * TY + cj) * (TX + 1) + ci;
					real a = real(0) - (fx[1] * aovX - fx[0] * aovX);
					if (DIM == 3) {
						real const* fy = fxy + (q * (TY + 1) + cj) * TX + ci;
						a = a - (fy[TX] * aovY - fy[0] * aovY);
					}
					accP[q] = a;
				}
			}
		}
	}
	if (sp.dtMinBits) {
		if (rateCell > real(0)) dtCell = rmin<real>(dtCell, real(1.) / rateCell);
		double v = double(dtCell);
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) { double const u = __shfl_xor_sync(0xffffffffu, v, o); v = u < v ? u : v; }
		if (lane == 0) redBuf[w] = v;
		__syncthreads();
		if (tid < 32) {
			v = tid < G::NWARPS ? redBuf[tid] : HUGE_VAL;
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) { double const u = __shfl_xor_sync(0xffffffffu, v, o); v = u < v ? u : v; }
			if (tid == 0 && v < HUGE_VAL) atomicMin(sp.dtMinBits, dtBits(v));
		}
	}
}

}   // namespace hb
