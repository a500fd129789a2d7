// hb_fv_march3.cuh -- second-generation plane-marching fused stage kernel for 3-D grids (sm_100a; TMA ring + split mbarrier).
//
// Same contract as fv_march (hb_fv_march.cuh): one launch = one Runge-Kutta stage =
//     calcLR ('plm cons', hydro/solver/plm.cl:32-91) + calcFlux (Roe, hydro/solver/fvsolver.lua:57-198, hydro/flux/roe.cl:17-163)
//     + calcDerivFromFlux (hydro/solver/fvsolver.cl:6-125) + the stage's multAdd combination (hydro/int/rk.lua:91-112)
//     + constrainU (hydro/solver/solverbase.lua:2116-2127) [+ calcDT / reduceMin on the last stage, hydro/eqn/cl/calcDT.cl:38-73]
// and the same arithmetic per cell (the -fmad=false build is bit-identical to fv_march, fv_stage and the CPU oracle).
//
// What round 1's ncu captures said about fv_march (profiles/r01c_fv_march_c4_*.txt): the kernel is bound by instruction issue and
// FP64 latency at 11 warps per SM, of which three (the halo warps) idle behind the per-plane __syncthreads while holding a
// quarter of the register file; 21 % of the stall samples are that barrier, and the shared-memory pipe (140 LDS/STS per cell)
// runs at 3/4 of the FP64 pipe's time.  This kernel changes the organisation, not the arithmetic:
//
//   * No slope exchange.  A thread forms BOTH face states of the interface it owns from the four cells of the stencil it reads
//     from the ring (slopes of cell i-1 and cell i share the middle difference): one more limiter per variable and side than
//     with an exchange, but the half-slope arrays, their halo rows/columns, the slope phase and its ordering constraint are gone.
//   * One halo warp instead of three: it computes the y fluxes of row TY (32 lanes) and the x fluxes of column TX (TY lanes), about
//     2/3 of a column warp's work, so every warp of the CTA is busy; TY = 15 gives 16 warps at 128 registers.
//   * Split barrier.  The only cross-warp hand-over is "low-face flux of plane k published" -> "high-face flux consumed".  Each warp
//     ARRIVES on an mbarrier after publishing and WAITS on it only after the marching-axis work of the same iteration (slope of
//     plane k+1, Roe flux at k+1/2: registers and the thread's own ring column only), so a warp that is ahead does a third of
//     an iteration's work instead of stalling.  The halo warp, which has no marching-axis work, waits at once and issues the TMA
//     request for plane k+3: the ring is refilled as soon as the last warp has left plane k-1.
//   * Cell k is finished inside iteration k (both z fluxes are known by then): no x/y partial sums and no copy of U[k-1] are
//     carried across iterations; the persistent registers are the face state and the flux of the previous z interface only.
//   * Production arithmetic: minmod by one FP64 compare + an integer sign test (hb_roe_fast.cuh: plmFacesT), fluxes accumulated
//     into the divergence as they are produced.
//   * RK operands (the other states / derivatives the stage's combination reads at the finished cell) come by TMA a plane ahead
//     (March3Cfg::OPTMA), and for classic-RK4-type tableaux the host carries the last stage's sum from stage to stage (StageP::Aout,
//     hb_fv.cu foldFinalStage), so every stage has at most two operands and runs the tallest tile.
#pragma once
#include "hb_fv_march.cuh"

namespace hb {

template<int TY_, int KM_, int VAR_ = 0> struct March3Cfg {
	static constexpr int TY = TY_;     // rows per CTA = column warps
	static constexpr int KM = KM_;     // planes per CTA along the marching axis
	static constexpr bool GRAV = (VAR_ & 32) != 0;      // the epilogue adds the self-gravity source (as MarchCfg::GRAV)
	// PAIR: the x and y flux cores of a cell issued as one block (two independent dependent chains).  MEASURED AND NOT USED: it spills at
	// 16 warps x 128 registers, changes nothing at 12 x 168 and gains 6 % at 8 x 208, where the kernel is 20 % slower anyway (r02w)
	static constexpr bool PAIR = (VAR_ & 1) != 0;
	// GEN: the general configuration -- any of the 20 slope limiters, no reconstruction, the Roe flux with a flux limiter (4-cell stencil
	// along every axis; hydro/solver/fvsolver.lua:138-155), and the other fluxes of the calcFluxForInterface slot (HLL, Rusanov, euler-HLLC),
	// all selected at run time from StageP as in the tile kernel fv_stage, with the literal device functions
	static constexpr bool GEN = (VAR_ & 2) != 0;
	// OPDIRECT: the RK operands are read from global memory in the epilogue (plain coalesced loads at the point of use) instead of being
	// staged per thread through shared memory by cp.async: no operand area in shared memory (every stage fits the tallest tile), no LDGSTS
	static constexpr bool OPDIRECT = (VAR_ & 8) != 0;
	// OPTMA (the default configurations): the RK operands of plane k arrive by TMA -- one cp.async.bulk.tensor per operand, issued by the halo
	// warp at the TOP of iteration k, as soon as the `opfree` mbarrier says every column warp has finished the epilogue of plane k-1; the
	// column warps wait on `opbar` just before their epilogue -- instead of 5 per-thread cp.async per operand: the LSU / MIO queue carries no
	// LDGSTS (profiles/r02k_fv_march3_c4_full.txt: RK4's four-operand stage stalled on mio_throttle + long_scoreboard).  2.18 -> 1.99 ms per
	// 512 x 512 x 128 stage; issuing the copies only when the iteration's flux barrier completes was too late to hide (2.36 ms)
	static constexpr bool OPTMA = (VAR_ & 16) != 0;
	// MINB 2 (VAR bit 6): two CTAs per SM (half-height tiles, registers capped for 2 x NT threads).  The idea: the warps of one CTA move
	// through the FP64-dense (flux cores) and FP64-sparse (ring loads, slopes, epilogue) phases of a plane together, a second, unsynchronised
	// CTA would fill the sparse phases.  MEASURED AND NOT USED: 2 x 8 warps run 2.03 ms against 1.93 ms for one 16-warp CTA (7 % more halo
	// work for 2 % better per row: the warps are not phase-locked, there are too few of them; profiles/r02n_sweep_minb2.txt)
	static constexpr int MINB = (VAR_ & 64) ? 2 : 1;
};

template<class C, class real> struct March3Geom {
	static constexpr int TX = 32, TY = C::TY;
	static constexpr int HL = (16 / int(sizeof(real))) > HB_G ? (16 / int(sizeof(real))) : HB_G;   // x halo of the TMA box (16-byte aligned start)
	static constexpr int BX = TX + 2 * HL, BY = TY + 2 * HB_G;
	static constexpr int PS = BX * BY;
	static constexpr int NCOL = TY, NWARPS = TY + 1, NT = 32 * NWARPS;
	static constexpr int R = 4;                                // ring slots: planes k, k+1, k+2 in use, k+3 in flight
	static constexpr int FXXN = TY * (TX + 1), FXYN = (TY + 1) * TX;
	template<int nI> static constexpr size_t slotBytes() { return (sizeof(real) * nI * PS + 127) / 128 * 128; }
	template<int nI> static constexpr size_t smemBytes(int nOps) {
		return 128 + R * slotBytes<nI>() + sizeof(real) * nI * size_t(2 * (FXXN + FXYN) + nOps * NCOL * 32) + 128;
	}
	static_assert(TY <= 31, "the halo warp's lanes serve the rows of column TX");
};

HB_D void mbarArrive(uint64_t* bar) {
	asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" :: "r"(smemAddr(bar)) : "memory");
}
// wait of a warp that is not on the critical path (the halo warp waiting for the column warps): back off between tests so that the
// polling does not take issue slots from the working warps
HB_D void mbarWaitBackoff(uint64_t* bar, uint32_t parity) {
	uint32_t done = 0;
	while (true) {
		asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(done) : "r"(smemAddr(bar)), "r"(parity) : "memory");
		if (done) break;
		__nanosleep(256);
	}
}

// Low-face Roe flux of the cell at ring offset `o` along the axis with ring stride `st` (1: x, BX: y): face states from the
// four-cell stencil o-2st .. o+st ('plm cons', plm.cl:56-76: UL = U[i-1] + .5 sigma[i-1], UR = U[i] - .5 sigma[i]).
// GEN: reconstruction / flux / limiter chosen at run time (see March3Cfg::GEN): mode 0 no reconstruction, 1 'plm cons', 2 flux limiter.
template<class Eqn, int SIDE, int LIM, int PS, bool GEN = false>
HB_D void lowFaceFlux(typename Eqn::real (&F)[Eqn::nI], typename Eqn::Params const& ep, int lim,
	typename Eqn::real const* __restrict__ P, int o, int st, int mode = 1, int flux = 0, int fluxParam = 0, int fluxLimiter = 0,
	typename Eqn::real dt_dx = 0)
{
	typedef typename Eqn::real real;
	constexpr int nI = Eqn::nI;
	if constexpr (GEN) {
		if (mode == 2) {
			real U2L[nI], UL[nI], UR[nI], U2R[nI];
			#pragma unroll
			for (int q = 0; q < nI; ++q) { real const* u = P + q * PS + o; U2L[q] = u[-2 * st]; UL[q] = u[-st]; UR[q] = u[0]; U2R[q] = u[st]; }
			roeFluxLimited<Eqn, SIDE>(F, ep, fluxLimiter, dt_dx, U2L, UL, UR, U2R);
			return;
		}
		real UL[nI], UR[nI];
		#pragma unroll
		for (int q = 0; q < nI; ++q) {
			real const* u = P + q * PS + o;
			if (mode == 1) { UL[q] = u[-st] + plmHalfSlope<real>(lim, u[-2 * st], u[-st], u[0]); UR[q] = u[0] - plmHalfSlope<real>(lim, u[-st], u[0], u[st]); }
			else { UL[q] = u[-st]; UR[q] = u[0]; }
		}
		interfaceFlux<Eqn, SIDE>(flux, fluxParam, F, ep, UL, UR);
	} else {
		real UL[nI], UR[nI];
		#pragma unroll
		for (int q = 0; q < nI; ++q) {
			real const* u = P + q * PS + o;
			plmFacesT<real, LIM, Eqn::FAST>(lim, u[-2 * st], u[-st], u[0], u[st], UL[q], UR[q]);
		}
		roeFluxAuto<Eqn, SIDE, true>(F, ep, UL, UR);
	}
}

// RK combination + constrainU + stores + CFL dt of one finished cell (hydro/int/rk.lua:96-112, solverbase.lua:2116-2127): the same
// operations in the same order as stageEpilogue (hb_fv_march.cuh), looping over the stage's compact term list (StageP::nT) instead of
// testing every slot of the alpha / beta tables.
template<class Eqn, bool GRAV, bool DIRECT = false>
HB_D void stageEpilogue3(GridP<typename Eqn::real> const& g, StageP<typename Eqn::real> const& sp, typename Eqn::Params const& ep,
	long long idx, typename Eqn::real (&acc)[Eqn::nI], typename Eqn::real const (&own)[Eqn::nI], double dt,
	typename Eqn::real& dtCell, typename Eqn::real& rateCell, typename Eqn::real const* ops, int opStride)
{
	typedef typename Eqn::real real;
	constexpr int nI = Eqn::nI;
	if constexpr (GRAV && Eqn::eqnId <= 1) {
		// op:addSource of the self-gravity op (solverbase.lua:3219-3223, selfgrav.cl:11-76), as in fv_stage
		if (sp.gravPot && sp.computeL) {
			real accel[3] = {0, 0, 0};
			accel[0] = (sp.gravPot[idx + 1] - sp.gravPot[idx - 1]) / (real(2.) * g.dx[0]);
			if (g.dim >= 2) accel[1] = (sp.gravPot[idx + g.strideY] - sp.gravPot[idx - g.strideY]) / (real(2.) * g.dx[1]);
			if (g.dim >= 3) accel[2] = (sp.gravPot[idx + g.strideZ] - sp.gravPot[idx - g.strideZ]) / (real(2.) * g.dx[2]);
			acc[1] = acc[1] - accel[0] * own[0];
			acc[2] = acc[2] - accel[1] * own[0];
			acc[3] = acc[3] - accel[2] * own[0];
			acc[4] -= own[1] * accel[0] + own[2] * accel[1] + own[3] * accel[2];
		}
	}
	if (sp.Lout) {
		#pragma unroll
		for (int q = 0; q < nI; ++q) sp.Lout[idx + q * g.strideV] = acc[q];
	}
	if (!sp.Uout) return;
	if (sp.Aout) {
		// the last stage's running sum (StageP::Aout): partial_s = partial_(s-1) + (beta_s dt) L_s, partial_(-1) = 0 + alpha U^0
		real const cb = real(sp.accBetaSelf * dt);
		if (sp.accSlot < 0) {
			real const ca = real(sp.accCoef);
			#pragma unroll
			for (int q = 0; q < nI; ++q) {
				real a = real(0) + own[q] * ca;
				a = a + acc[q] * cb;
				sp.Aout[idx + q * g.strideV] = a;
			}
		} else {
			#pragma unroll
			for (int q = 0; q < nI; ++q) {
				real a;
				if constexpr (DIRECT) a = __ldg(sp.opPtr[sp.accSlot] + idx + q * g.strideV);
				else a = ops[(sp.accSlot * nI + q) * opStride];
				a = a + acc[q] * cb;
				sp.Aout[idx + q * g.strideV] = a;
			}
		}
	}
	real U[nI];
	#pragma unroll
	for (int q = 0; q < nI; ++q) U[q] = 0;
	#pragma unroll 1
	for (int t = 0; t < sp.nT; ++t) {
		double cd = sp.tCoef[t];
		if ((sp.tBetaMask >> t) & 1) cd = cd * dt;
		real const c = real(cd);
		int const slot = sp.tSlot[t];
		if (slot < 0) {
			#pragma unroll
			for (int q = 0; q < nI; ++q) U[q] = U[q] + own[q] * c;
		} else {
			if constexpr (DIRECT) {
				real const* o = sp.opPtr[slot] + idx;
				real v[nI];
				#pragma unroll
				for (int q = 0; q < nI; ++q) v[q] = __ldg(o + q * g.strideV);
				#pragma unroll
				for (int q = 0; q < nI; ++q) U[q] = U[q] + v[q] * c;
			} else {
				real const* o = ops + slot * nI * opStride;
				#pragma unroll
				for (int q = 0; q < nI; ++q) U[q] = U[q] + o[q * opStride] * c;
			}
		}
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

template<class Eqn, int LIM, class C, int MODE>
__global__ void __launch_bounds__((March3Geom<C, typename Eqn::real>::NT), C::MINB)
fv_march3(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ GridP<typename Eqn::real> g,
	const __grid_constant__ StageP<typename Eqn::real> sp, const __grid_constant__ typename Eqn::Params ep, int const padX, int const chunkSel)
{
	typedef typename Eqn::real real;
	typedef March3Geom<C, real> G;
	constexpr int nI = Eqn::nI;
	constexpr bool FAST = Eqn::FAST;
	constexpr int TX = G::TX, TY = G::TY, BX = G::BX, PS = G::PS;
	constexpr int SLOT = (int(sizeof(real)) * nI * PS + 127) / 128 * 128 / int(sizeof(real));   // == slotBytes / sizeof(real)
	extern __shared__ __align__(128) unsigned char march3Smem[];
	uint64_t* full = reinterpret_cast<uint64_t*>(march3Smem);              // R mbarriers: ring slot filled
	uint64_t* xbar = full + G::R;                                          // 2 mbarriers: fluxes of plane k published (by iteration parity)
	uint64_t* opbar = xbar + 2;                                            // OPTMA: the RK operands of plane k have landed
	uint64_t* opfree = opbar + 1;                                          // OPTMA: every column warp has finished the epilogue of plane k (operand area free)
	real* ring = reinterpret_cast<real*>(march3Smem + 128);
	real* FXX = ring + G::R * SLOT;             // x fluxes at the low faces of cells i = 0 .. TX  [parity][q][row][i]
	real* FXY = FXX + 2 * nI * G::FXXN;         // y fluxes at the low faces of rows j = 0 .. TY   [parity][q][j][i]
	// staged RK operands of the column threads [operand][q][thread], on a 128-byte boundary (a TMA destination when C::OPTMA; the size formula
	// carries the slack)
	real* OPB = reinterpret_cast<real*>(march3Smem + ((reinterpret_cast<unsigned char*>(FXY + 2 * nI * G::FXYN) - march3Smem) + 127) / 128 * 128);
	constexpr int OPS = G::NCOL * 32;
	__shared__ double redBuf[32];

	int const tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
	int const lim = LIM >= 0 ? LIM : sp.slopeLimiter;
	bool const isCol = w < TY;
	constexpr bool GEN = C::GEN;
	// GEN: 0 no reconstruction, 1 'plm cons', 2 Roe with a flux limiter on cell-centred states (gridsolver.lua:119: never both)
	int const gmode = !GEN ? 1 : (sp.plmMode == 1 ? 1 : (sp.fluxLimiter > 0 ? 2 : 0));
	real const dtR = GEN ? real(*sp.dt) : real(0);

	// ---- tile
	int const ntx = (g.N[0] + TX - 1) / TX;
	int const nty = (g.N[1] + TY - 1) / TY;
	int bid = blockIdx.x;
	int const bx = bid % ntx; bid /= ntx;
	int const by = bid % nty; int bm = bid / nty;
	int const i0 = bx * TX + HB_G, j0 = by * TY + HB_G;
	// chunkSel (overlapped slab exchange, hb_fv.cu): 0 = every plane, in chunks of KM; 1 = the RIM: the HB_G lowest (bm = 0) and the HB_G
	// highest (bm = 1) interior planes, whose values the neighbouring slabs need; 2 = the planes in between, in chunks of KM
	int kb, ke;
	if (chunkSel == 1) { kb = bm == 0 ? HB_G : g.N[2]; ke = kb + HB_G; }
	else if (chunkSel == 2) { kb = 2 * HB_G + bm * C::KM; ke = min(kb + C::KM, g.N[2]); }
	else { kb = HB_G + bm * C::KM; ke = min(kb + C::KM, HB_G + g.N[2]); }      // ke exclusive

	uint32_t const boxBytes = uint32_t(sizeof(real) * nI * PS);
	int const tx0 = i0 - G::HL + padX, ty0 = j0 - HB_G;
	auto issue = [&](int plane) {
		int const slot = (plane - (kb - 2)) & 3;
		mbarExpectTx(&full[slot], boxBytes);
		tmaLoad4D(ring + slot * SLOT, &tmap, &full[slot], tx0, ty0, plane, 0);
	};
	auto waitPlane = [&](int plane) {
		int const n = plane - (kb - 2);
		mbarWait(&full[n & 3], uint32_t(n >> 2) & 1u);
	};
	if (tid == 0) {
		for (int s = 0; s < G::R; ++s) mbarInit(&full[s], 1);
		mbarInit(&xbar[0], G::NWARPS);
		mbarInit(&xbar[1], G::NWARPS);
		mbarInit(opbar, 1);
		mbarInit(opfree, G::NCOL);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads();
	if (tid == 0) { issue(kb - 2); issue(kb - 1); issue(kb); issue(kb + 1); }

	real dtCell = inf_of<real>::v(), rateCell = 0;

	if (!isCol) {
		// ================= halo warp: y fluxes of row TY, x fluxes of column TX, ring refill =================
		int const oY = (TY + HB_G) * BX + lane + G::HL;                    // cell (lane, TY)
		int const rowX = lane < TY ? lane : TY - 1;
		int const oX = (rowX + HB_G) * BX + TX + G::HL;                    // cell (TX, lane)
		bool const on0 = FAST || g.fluxOn[0], on1 = FAST || g.fluxOn[1];   // (production: a switched-off side has aov = 0 at the consumer)
		for (int k = kb - 1, it = 0; k < ke; ++k, ++it) {
			if constexpr (C::OPTMA) {
				// the RK operands of plane k, consumed at the END of this iteration: fetched now, a whole iteration ahead, as soon as every column
				// warp has finished the epilogue of plane k-1 (the operand area's last reader)
				if (k >= kb && sp.nOps > 0 && sp.Uout) {
					if (k > kb) mbarWaitBackoff(opfree, uint32_t(k - 1 - kb) & 1u);
					if (lane == 0) {
						asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
						mbarExpectTx(opbar, uint32_t(sizeof(real) * nI * OPS) * uint32_t(sp.nOps));
						for (int o = 0; o < sp.nOps; ++o)
							tmaLoad4D(OPB + o * (nI * OPS), reinterpret_cast<const CUtensorMap*>(sp.opMaps) + o, opbar, i0 + padX, j0, k, 0);
					}
				}
			}
			if (k >= kb) {
				waitPlane(k);
				real const* __restrict__ P = ring + ((k - (kb - 2)) & 3) * SLOT;
				real* const fxx = FXX + (k & 1) * (nI * G::FXXN);
				real* const fxy = FXY + (k & 1) * (nI * G::FXYN);
				real F[nI];
				lowFaceFlux<Eqn, 1, LIM, PS, GEN>(F, ep, lim, P, oY, BX, gmode, sp.flux, sp.fluxParam, sp.fluxLimiter, GEN ? dtR / g.dx[1] : real(0));
				#pragma unroll
				for (int q = 0; q < nI; ++q) fxy[(q * (TY + 1) + TY) * TX + lane] = on1 ? F[q] : real(0);
				lowFaceFlux<Eqn, 0, LIM, PS, GEN>(F, ep, lim, P, oX, 1, gmode, sp.flux, sp.fluxParam, sp.fluxLimiter, GEN ? dtR / g.dx[0] : real(0));
				if (lane < TY) {
					#pragma unroll
					for (int q = 0; q < nI; ++q) fxx[(q * TY + lane) * (TX + 1) + TX] = on0 ? F[q] : real(0);
				}
			}
			__syncwarp();
			if (lane == 0) mbarArrive(&xbar[it & 1]);
			mbarWaitBackoff(&xbar[it & 1], uint32_t(it >> 1) & 1u);
			// every warp has left plane k-1 (its last reader is the epilogue of iteration k-1): its slot takes plane k+3
			if (lane == 0 && k + 3 <= ke + 1) {
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
				issue(k + 3);
			}
		}
	} else {
		// ================= column warps =================
		int const cj = w, ci = lane;
		int const ob = (cj + HB_G) * BX + (ci + G::HL);                    // own position inside a slot (per variable)
		int const gi = i0 + ci, gj = j0 + cj;
		bool const inside = gi < g.S[0] - HB_G && gj < g.S[1] - HB_G;
		long long const colIdx = gi + g.strideY * gj;
		long long const strideM = g.strideZ;
		double const dt = *sp.dt;
		real const aovX = g.aov[0], aovY = g.aov[1], aovM = g.aov[2];

		// face state of plane kb-1 towards kb (slope of plane kb-1 from planes kb-2, kb-1, kb)
		real zf[nI];
		real zacc[nI];          // production: F_z(k-1/2) aov_z, the start of cell k's divergence sum; literal: F_z(k-1/2)
		waitPlane(kb - 2); waitPlane(kb - 1); waitPlane(kb);
		#pragma unroll
		for (int q = 0; q < nI; ++q) {
			real const a = ring[0 * SLOT + q * PS + ob], b = ring[1 * SLOT + q * PS + ob], c = ring[2 * SLOT + q * PS + ob];
			if constexpr (GEN) {
				// zf: mode 1 the face state of plane kb-1 towards kb; mode 2 the cell of plane kb-2 (U2L of the first z interface); mode 0 unused
				zf[q] = gmode == 1 ? b + plmHalfSlope<real>(lim, a, b, c) : a;
			} else {
				real lo;
				plmCellFacesT<real, LIM, FAST>(lim, a, b, c, lo, zf[q]);
			}
			zacc[q] = 0;
		}

		for (int k = kb - 1, it = 0; k < ke; ++k, ++it) {
			bool const own = k >= kb;                                      // cells of plane k are finished in this iteration
			real const* __restrict__ P = ring + ((k - (kb - 2)) & 3) * SLOT;
			real* const fxx = FXX + (k & 1) * (nI * G::FXXN);
			real* const fxy = FXY + (k & 1) * (nI * G::FXYN);
			long long const idxK = colIdx + strideM * k;
			real acc[nI];
			// ---- x and y: Roe fluxes at the low faces of the own cell, published for the neighbours towards -x / -y
			if (own) {
				if (!C::OPDIRECT && !C::OPTMA && inside && sp.Uout) {
					// the RK operands of cell k are consumed at the end of this iteration: start their global -> shared copies now
					// (per-thread slots: no barrier involved)
					#pragma unroll 1
					for (int o = 0; o < sp.nOps; ++o) {
						real const* src = sp.opPtr[o] + idxK;
						real* dst = OPB + o * (nI * OPS) + tid;
						#pragma unroll
						for (int q = 0; q < nI; ++q) cpAsyncElem<real>(dst + q * OPS, src + q * g.strideV);
					}
					cpAsyncCommit();
				}
				if constexpr (C::PAIR && FAST) {
					real ULx[nI], URx[nI], ULy[nI], URy[nI], Fx[nI], Fy[nI];
					#pragma unroll
					for (int q = 0; q < nI; ++q) {
						real const* u = P + q * PS + ob;
						plmFacesT<real, LIM, FAST>(lim, u[-2], u[-1], u[0], u[1], ULx[q], URx[q]);
						plmFacesT<real, LIM, FAST>(lim, u[-2 * BX], u[-BX], u[0], u[BX], ULy[q], URy[q]);
					}
					roeFluxPairAuto<Eqn, 0, 1>(Fx, Fy, ep, ULx, URx, ULy, URy);
					#pragma unroll
					for (int q = 0; q < nI; ++q) {
						acc[q] = fma(Fy[q], aovY, fma(Fx[q], aovX, zacc[q]));
						fxx[(q * TY + cj) * (TX + 1) + ci] = Fx[q];
						fxy[(q * (TY + 1) + cj) * TX + ci] = Fy[q];
					}
				} else {
				real F[nI];
				lowFaceFlux<Eqn, 0, LIM, PS, GEN>(F, ep, lim, P, ob, 1, gmode, sp.flux, sp.fluxParam, sp.fluxLimiter, GEN ? dtR / g.dx[0] : real(0));
				#pragma unroll
				for (int q = 0; q < nI; ++q) {
					if constexpr (FAST) acc[q] = fma(F[q], aovX, zacc[q]);
					else if (!g.fluxOn[0]) F[q] = 0;
					fxx[(q * TY + cj) * (TX + 1) + ci] = F[q];
				}
				lowFaceFlux<Eqn, 1, LIM, PS, GEN>(F, ep, lim, P, ob, BX, gmode, sp.flux, sp.fluxParam, sp.fluxLimiter, GEN ? dtR / g.dx[1] : real(0));
				#pragma unroll
				for (int q = 0; q < nI; ++q) {
					if constexpr (FAST) acc[q] = fma(F[q], aovY, acc[q]);
					else if (!g.fluxOn[1]) F[q] = 0;
					fxy[(q * (TY + 1) + cj) * TX + ci] = F[q];
				}
				}
			}
			__syncwarp();
			if (lane == 0) mbarArrive(&xbar[it & 1]);
			// ---- marching axis (registers + the thread's own ring column): slope of plane k+1, Roe flux at k+1/2
			waitPlane(k + 2);
			real Fz[nI];
			{
				real const* __restrict__ Q1 = ring + ((k + 1 - (kb - 2)) & 3) * SLOT;
				real const* __restrict__ Q2 = ring + ((k + 2 - (kb - 2)) & 3) * SLOT;
				if constexpr (GEN) {
					real UL[nI], UR[nI], U2R[nI];
					#pragma unroll
					for (int q = 0; q < nI; ++q) { UL[q] = P[q * PS + ob]; UR[q] = Q1[q * PS + ob]; U2R[q] = Q2[q * PS + ob]; }
					if (gmode == 2) {
						roeFluxLimited<Eqn, 2>(Fz, ep, sp.fluxLimiter, dtR / g.dx[2], zf, UL, UR, U2R);
						#pragma unroll
						for (int q = 0; q < nI; ++q) zf[q] = UL[q];
					} else if (gmode == 1) {
						real URf[nI];
						#pragma unroll
						for (int q = 0; q < nI; ++q) {
							real const sN = plmHalfSlope<real>(lim, UL[q], UR[q], U2R[q]);
							URf[q] = UR[q] - sN;
							U2R[q] = UR[q] + sN;                    // (reused: the face state of plane k+1 towards k+2)
						}
						interfaceFlux<Eqn, 2>(sp.flux, sp.fluxParam, Fz, ep, zf, URf);
						#pragma unroll
						for (int q = 0; q < nI; ++q) zf[q] = U2R[q];
					} else interfaceFlux<Eqn, 2>(sp.flux, sp.fluxParam, Fz, ep, UL, UR);
				} else {
				real UR[nI], zfN[nI];
				#pragma unroll
				for (int q = 0; q < nI; ++q) plmCellFacesT<real, LIM, FAST>(lim, P[q * PS + ob], Q1[q * PS + ob], Q2[q * PS + ob], UR[q], zfN[q]);
				roeFluxAuto<Eqn, 2, true>(Fz, ep, zf, UR);
				#pragma unroll
				for (int q = 0; q < nI; ++q) zf[q] = zfN[q];
				}
			}
			if constexpr (FAST) {
				#pragma unroll
				for (int q = 0; q < nI; ++q) {
					real const t = Fz[q] * aovM;
					acc[q] = acc[q] - t;
					zacc[q] = t;
				}
			} else if (!g.fluxOn[2]) {
				#pragma unroll
				for (int q = 0; q < nI; ++q) Fz[q] = 0;
			}
			mbarWait(&xbar[it & 1], uint32_t(it >> 1) & 1u);
			if constexpr (C::OPTMA) {
				if (own && sp.nOps > 0 && sp.Uout) mbarWait(opbar, uint32_t(k - kb) & 1u);
			}
			// ---- flux differences (fvsolver.cl:97-123) and the epilogue of cell k
			if (own && inside) {
				real U0[nI];
				#pragma unroll
				for (int q = 0; q < nI; ++q) {
					real const* fx = fxx + (q * TY + cj) * (TX + 1) + ci;
					real const* fy = fxy + (q * (TY + 1) + cj) * TX + ci;
					if constexpr (FAST) {
						acc[q] = fma(-fy[TX], aovY, fma(-fx[1], aovX, acc[q]));
					} else {
						real a = real(0) - (fx[1] * aovX - fx[0] * aovX);
						a = a - (fy[TX] * aovY - fy[0] * aovY);
						acc[q] = g.volOn ? a - (Fz[q] * aovM - zacc[q] * aovM) : real(0);
					}
					U0[q] = P[q * PS + ob];
				}
				if constexpr (!C::OPDIRECT && !C::OPTMA) cpAsyncWaitAll();
				stageEpilogue3<Eqn, C::GRAV, C::OPDIRECT>(g, sp, ep, idxK, acc, U0, dt, dtCell, rateCell, OPB + tid, OPS);
			}
			if constexpr (C::OPTMA) {
				if (own && sp.nOps > 0 && sp.Uout) { __syncwarp(); if (lane == 0) mbarArrive(opfree); }
			}
			if constexpr (!FAST) {
				#pragma unroll
				for (int q = 0; q < nI; ++q) zacc[q] = Fz[q];
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
