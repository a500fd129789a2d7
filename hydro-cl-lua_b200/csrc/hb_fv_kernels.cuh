// hb_fv_kernels.cuh -- the fused finite-volume stage kernel and its helper kernels (sm_100a).
//
// One launch of `fv_stage` replaces, for one Runge-Kutta stage, this sequence of reference kernels
// (hydro/solver/fvsolver.lua:225-302, hydro/int/rk.lua:91-165, hydro/solver/solverbase.lua:2116-2127):
//     fill(derivBuf,0); calcLR; calcFlux; calcDerivFromFlux; fill(UBuf,0); multAdd x (#alpha + #beta);
//     constrainU            [+ calcDT, reduceMin on the last stage]
// The ghost-cell fill (boundary_x/y/z, gridsolver.lua:1070-1213) is `fill_ghosts`, one launch per stage.
//
// Data layout: structure of arrays  U[var][k][j][i]  (i fastest), ghost cells kept in the array so that
// cell indices equal the reference's INDEX macro (hydro/app.lua:976-984).  Only the nI integrated
// variables are touched by a stage; the remaining cons_t fields live in the same allocation.
//
// Tile algorithm (per CTA, tile = TX x TY x TZ interior cells):
//   0. stage the tile plus a 2-cell halo of every integrated variable in shared memory
//   for each side s:
//     A. [PLM] half slopes .5*sigma_s of every cell of the tile extended by one cell along s  -> smem SG
//     B. Roe flux at the T_s+1 interfaces along s of every pencil                              -> smem FX
//     C. each thread subtracts (F_hi - F_lo) * area/volume for the cells it owns (registers)
//   E. epilogue per owned cell: RK combination (alpha terms, then beta terms, k ascending -- rk.lua:96-112),
//      constrainU, store; optional store of L(U); optional CFL dt min -> warp shuffle -> block -> atomicMin.
// Work in A and B is laid out so that a warp's lanes walk a plane perpendicular to s (whole warps fall on
// one layer), which keeps every lane busy although the halo layers need slopes but no flux.
#pragma once
#include "hb_math.cuh"
#include "hb_roe.cuh"
#include "hb_eqn_euler.cuh"
#include "hb_eqn_mhd.cuh"

namespace hb {

constexpr int HB_MAX_TERMS = 4;
constexpr int HB_G = 2;   // numGhost, hydro/solver/gridsolver.lua:41

template<class real> struct GridP {
	int dim;
	int S[3];              // ghost-inclusive size per axis (1 on unused axes), gridsolver.lua:94-95
	int N[3];              // interior size
	long long strideY, strideZ, strideV;   // element strides of j, k and of the variable index
	real dx[3];            // gridsolver.lua:406-409
	real invdx[3];         // 1 / dx (production CFL reduction)
	real aov[3];           // area_s * (1/volume) as calcDerivFromFlux forms it (fvsolver.cl:97-102)
	int fluxOn[3];         // area_s > 1e-7 (fvsolver.lua:107-110)
	int volOn;             // volume > 1e-7 (fvsolver.cl:97)
};

template<class real> struct StageP {
	const real* Uin;       // stage input state (ghosts valid)
	real* Uout;            // next stage state (interior written)
	real* Lout;            // optional: dU/dt of Uin (interior), for later stages' beta terms
	int nA; const real* aPtr[HB_MAX_TERMS]; double aCoef[HB_MAX_TERMS];   // alpha_k * U^k
	int aOwnMask;          // bit a set: aPtr[a] == Uin (the marching kernel then uses its register copy)
	int nB; const real* bPtr[HB_MAX_TERMS]; double bCoef[HB_MAX_TERMS];   // (beta_k dt) * L^k, k < this stage
	double betaSelf;       // beta of this stage's own L; used when computeL
	int computeL;
	const double* dt;      // device scalar: the full-step dt (solverbase.lua:3197-3202: stages see the step dt)
	unsigned long long* dtMinBits;   // optional: fused calcDT, min over interior cells as ordered bits of a double
	int slopeLimiter, fluxLimiter;
	real* scratch;         // optional per-solver device scratch (FvOps::scratchElems), e.g. the ADM flux arrays
	int flux;              // HB_FLUX_*: 0 roe, 1 hll, 2 rusanov, 3 euler-hllc (tile kernel; the marching kernel is built for roe)
	int fluxParam;         // euler-hllc: hllcMethod
	const real* gravPot;   // optional: the potential (ePot of Uin) of the self-gravity op; the tile kernel adds calcGravityDeriv (selfgrav.cl:53-76) to L
	int plmMode;           // hb_fv_desc.use_plm: 0 none, 1 'plm cons', 2 'plm athena' (faces as the reference tree assigns them), 3 'plm athena' with L/R as recorded, 4 'plm prim', 5 'plm cons with flux', 6 'plm eig', 7 'plm eig prim', 8 'plm eig prim ref', 9 / 10 = 7 / 8 with L and R exchanged
	// the same RK combination as a compact list in evaluation order (alpha terms, then beta terms; rk.lua:96-112), for kernels that loop
	// over it (fv_march3): term t adds tCoef[t] (x dt when bit t of tBetaMask is set) times the stage input (tSlot[t] < 0) or staged operand tSlot[t]
	int nT, tBetaMask, nOps;
	int tSlot[2 * HB_MAX_TERMS];
	double tCoef[2 * HB_MAX_TERMS];
	const real* opPtr[2 * HB_MAX_TERMS];
	// optional (fv_march3 configurations with TMA-staged operands): device array of nOps tensor maps (CUtensorMap, 128 bytes each) over the
	// operand buffers, box = {32, TY, 1, nI}: the halo warp fetches plane k of every operand with one bulk copy each instead of 5 nOps
	// per-thread cp.async
	const void* opMaps;
	// optional second output (marching kernels; hb_fv.cu foldFinalStage): the running sum of the LAST stage's combination,
	// Aout = (accSlot == -1 ? 0 + accCoef * stage input : staged operand accSlot) + (accBetaSelf dt) * L -- the partial sums of rk.lua:96-112 in
	// the reference's order, carried from stage to stage so that the last stage reads one operand instead of every earlier L
	real* Aout;
	int accSlot;
	double accCoef, accBetaSelf;
};

template<int TX_, int TY_, int TZ_, int NT_> struct Tile {
	static constexpr int TX = TX_, TY = TY_, TZ = TZ_, NT = NT_;
};

HB_D unsigned long long dtBits(double v) { return (unsigned long long)__double_as_longlong(v); }

template<int DIM, class T> struct TileGeom {
	static constexpr int TX = T::TX, TY = DIM >= 2 ? T::TY : 1, TZ = DIM >= 3 ? T::TZ : 1;
	static constexpr int GY = DIM >= 2 ? HB_G : 0, GZ = DIM >= 3 ? HB_G : 0;
	static constexpr int BX = TX + 2 * HB_G, BY = TY + 2 * GY, BZ = TZ + 2 * GZ;
	static constexpr int BXP = BX | 1;   // odd row pitch: column walks stay conflict-free
	static constexpr int BOX = BXP * BY * BZ;
	static constexpr int CELLS = TX * TY * TZ;
	static constexpr int P0 = TY * TZ, P1 = TX * TZ, P2 = TX * TY;   // pencils per side
	static constexpr int P0P = (P0 % 2 == 0 && P0 > 1) ? P0 + 1 : P0;   // padded layer pitch of side 0 in FX (phase C walks i)
	static constexpr int cmax(int a, int b) { return a > b ? a : b; }
	static constexpr int SGN = cmax((TX + 2) * P0, cmax(DIM >= 2 ? (TY + 2) * P1 : 0, DIM >= 3 ? (TZ + 2) * P2 : 0));
	static constexpr int FXN = cmax((TX + 1) * P0P, cmax(DIM >= 2 ? (TY + 1) * P1 : 0, DIM >= 3 ? (TZ + 1) * P2 : 0));
	static constexpr int CPT = (CELLS + T::NT - 1) / T::NT;
	// plm: 0 none, 1 'plm cons' (one half slope per variable), 2 'plm athena' (both face states per variable)
	template<class real, int nI> static constexpr size_t smemBytes(int plm) {
		return sizeof(real) * size_t(nI) * (BOX + plm * SGN + FXN) + 64;
	}
};

// box offset of tile-local cell (i,j,k), each coordinate in [-2, T+2)
template<int DIM, class T> HB_D int boxIdx(int i, int j, int k) {
	typedef TileGeom<DIM, T> G;
	return ((k + G::GZ) * G::BY + (j + G::GY)) * G::BXP + (i + HB_G);
}

// decode item w of side SIDE: layer c (position along SIDE) and the pencil's transverse coordinates
template<int DIM, class T, int SIDE> HB_D void decodeItem(int w, int& c, int& p, int& i, int& j, int& k) {
	typedef TileGeom<DIM, T> G;
	if (SIDE == 0) { c = w / G::P0; p = w - c * G::P0; j = p % G::TY; k = p / G::TY; i = c; }
	else if (SIDE == 1) { c = w / G::P1; p = w - c * G::P1; i = p % G::TX; k = p / G::TX; j = c; }
	else { c = w / G::P2; p = w - c * G::P2; i = p % G::TX; j = p / G::TX; k = c; }
}

template<class Eqn, int DIM, bool PLM, bool FLIM, class T, int SIDE>
HB_D void stageSide(GridP<typename Eqn::real> const& g, StageP<typename Eqn::real> const& sp,
	typename Eqn::Params const& ep, typename Eqn::real const* __restrict__ Us, typename Eqn::real* __restrict__ SG,
	typename Eqn::real* __restrict__ FX, typename Eqn::real (&acc)[TileGeom<DIM, T>::CPT][Eqn::nI], typename Eqn::real dtReal)
{
	typedef typename Eqn::real real;
	typedef TileGeom<DIM, T> G;
	constexpr int nI = Eqn::nI;
	constexpr int TS = SIDE == 0 ? G::TX : (SIDE == 1 ? G::TY : G::TZ);
	constexpr int P = SIDE == 0 ? G::P0 : (SIDE == 1 ? G::P1 : G::P2);
	constexpr int PF = SIDE == 0 ? G::P0P : P;   // layer pitch in FX
	constexpr int step = SIDE == 0 ? 1 : (SIDE == 1 ? G::BXP : G::BXP * G::BY);
	int const tid = threadIdx.x;

	if (PLM) {
		// ---- A: half slopes for cells c = -1 .. TS along SIDE
		for (int w = tid; w < (TS + 2) * P; w += T::NT) {
			int c, p, i, j, k;
			decodeItem<DIM, T, SIDE>(w, c, p, i, j, k);
			if (SIDE == 0) i -= 1; else if (SIDE == 1) j -= 1; else k -= 1;
			int const b = boxIdx<DIM, T>(i, j, k);
			if (sp.plmMode >= 2) {
				// 'plm athena': both face states of the cell -> SG (L faces), SG + nI * SGN (R faces)
				if constexpr (Eqn::hasEigenForCell) {
					real UL[nI], U[nI], UR[nI], L[nI], R[nI];
					#pragma unroll
					for (int q = 0; q < nI; ++q) {
						real const* u = Us + q * G::BOX + b;
						UL[q] = u[-step]; U[q] = u[0]; UR[q] = u[step];
					}
					if (sp.plmMode == 4) plmPrimFaces<Eqn>(L, R, ep, sp.slopeLimiter, UL, U, UR);
					else if (sp.plmMode == 5) plmConsFluxFaces<Eqn, SIDE>(L, R, ep, sp.slopeLimiter, dtReal / g.dx[SIDE], UL, U, UR);
					else if (sp.plmMode == 6) plmEigFaces<Eqn, SIDE>(L, R, ep, sp.slopeLimiter, dtReal / g.dx[SIDE], UL, U, UR);
					else if (sp.plmMode >= 7 && sp.plmMode <= 10) plmEigPrimFaces<Eqn, SIDE>(L, R, ep, sp.plmMode == 8 || sp.plmMode == 10, sp.plmMode >= 9 ? 1 : 0, dtReal / g.dx[SIDE], UL, U, UR);
					else plmAthenaFaces<Eqn, SIDE>(L, R, ep, UL, U, UR, sp.plmMode == 3 ? 1 : 0);
					#pragma unroll
					for (int q = 0; q < nI; ++q) { SG[q * G::SGN + w] = L[q]; SG[(nI + q) * G::SGN + w] = R[q]; }
				}
			} else {
				#pragma unroll
				for (int q = 0; q < nI; ++q) {
					real const* u = Us + q * G::BOX + b;
					SG[q * G::SGN + w] = plmHalfSlope<real>(sp.slopeLimiter, u[-step], u[0], u[step]);
				}
			}
		}
		__syncthreads();
	}
	if (!PLM && SIDE > 0) __syncthreads();   // phase C of the previous side still reads FX
	// ---- B: Roe flux at interfaces f = 0 .. TS (low face of cell f)
	for (int w = tid; w < (TS + 1) * P; w += T::NT) {
		int f, p, i, j, k;
		decodeItem<DIM, T, SIDE>(w, f, p, i, j, k);
		int const b = boxIdx<DIM, T>(i, j, k);   // cell f (right of the interface)
		real F[nI];
		if (!g.fluxOn[SIDE]) {
			#pragma unroll
			for (int q = 0; q < nI; ++q) F[q] = 0;
		} else if (FLIM) {
			real U2L[nI], UL[nI], UR[nI], U2R[nI];
			#pragma unroll
			for (int q = 0; q < nI; ++q) {
				real const* u = Us + q * G::BOX + b;
				U2L[q] = u[-2 * step]; UL[q] = u[-step]; UR[q] = u[0]; U2R[q] = u[step];
			}
			roeFluxLimited<Eqn, SIDE>(F, ep, sp.fluxLimiter, dtReal / g.dx[SIDE], U2L, UL, UR, U2R);
		} else {
			real UL[nI], UR[nI];
			#pragma unroll
			for (int q = 0; q < nI; ++q) {
				real const* u = Us + q * G::BOX + b;
				if (PLM && sp.plmMode >= 2) {
					UL[q] = SG[(nI + q) * G::SGN + w];            // R face state of cell f-1 (item layer f-1+1)
					UR[q] = SG[q * G::SGN + w + P];               // L face state of cell f
				} else if (PLM) {
					UL[q] = u[-step] + SG[q * G::SGN + w];        // R face state of cell f-1 (item layer f-1+1)
					UR[q] = u[0] - SG[q * G::SGN + w + P];        // L face state of cell f
				} else {
					UL[q] = u[-step]; UR[q] = u[0];
				}
			}
			interfaceFlux<Eqn, SIDE>(sp.flux, sp.fluxParam, F, ep, UL, UR);
		}
		#pragma unroll
		for (int q = 0; q < nI; ++q) FX[q * G::FXN + f * PF + p] = F[q];
	}
	__syncthreads();
	// ---- C: flux difference of the owned cells (fvsolver.cl:97-123)
	if (g.volOn) {
		real const aov = g.aov[SIDE];
		#pragma unroll
		for (int n = 0; n < G::CPT; ++n) {
			int const cell = tid + n * T::NT;
			if (G::CELLS % T::NT != 0 && cell >= G::CELLS) break;
			int const i = cell % G::TX, j = (cell / G::TX) % G::TY, k = cell / (G::TX * G::TY);
			int c, p;
			if (SIDE == 0) { c = i; p = j + G::TY * k; }
			else if (SIDE == 1) { c = j; p = i + G::TX * k; }
			else { c = k; p = i + G::TX * j; }
			#pragma unroll
			for (int q = 0; q < nI; ++q) {
				real const fl = FX[q * G::FXN + c * PF + p];
				real const fr = FX[q * G::FXN + (c + 1) * PF + p];
				acc[n][q] = acc[n][q] - (fr * aov - fl * aov);
			}
		}
	}
}

// MODE (0 = production, 1 = strict -fmad=false build) only makes the two builds' kernels distinct symbols: the
// translation units are linked into one library and must not share host stubs.
template<class Eqn, int DIM, bool PLM, bool FLIM, class T, int MODE>
__global__ void __launch_bounds__(T::NT)
fv_stage(GridP<typename Eqn::real> const g, StageP<typename Eqn::real> const sp, typename Eqn::Params const ep)
{
	typedef typename Eqn::real real;
	typedef TileGeom<DIM, T> G;
	constexpr int nI = Eqn::nI;
	extern __shared__ __align__(16) unsigned char smemRaw[];
	real* Us = reinterpret_cast<real*>(smemRaw);
	real* SG = Us + nI * G::BOX;
	real* FX = SG + (PLM ? (sp.plmMode >= 2 ? 2 : 1) * nI * G::SGN : 0);
	__shared__ double redBuf[32];

	int const tid = threadIdx.x;
	// tile origin (padded-array coordinates of tile-local cell 0)
	int const ntx = (g.N[0] + G::TX - 1) / G::TX;
	int const nty = (g.N[1] + G::TY - 1) / G::TY;
	int bid = blockIdx.x;
	int const bx = bid % ntx; bid /= ntx;
	int const by = bid % nty; int const bz = bid / nty;
	int const i0 = bx * G::TX + HB_G;
	int const j0 = DIM >= 2 ? by * G::TY + HB_G : 0;
	int const k0 = DIM >= 3 ? bz * G::TZ + HB_G : 0;

	// ---- 0: stage tile + halo (indices clamped into the array; clamped copies are never used for output)
	for (int w = tid; w < G::BX * G::BY * G::BZ; w += T::NT) {
		int const i = w % G::BX, j = (w / G::BX) % G::BY, k = w / (G::BX * G::BY);
		int gi = i0 - HB_G + i; gi = gi < g.S[0] ? gi : g.S[0] - 1;
		int gj = j0 - G::GY + j; gj = gj < g.S[1] ? gj : g.S[1] - 1;
		int gk = k0 - G::GZ + k; gk = gk < g.S[2] ? gk : g.S[2] - 1;
		long long const src = gi + g.strideY * gj + g.strideZ * gk;
		int const dst = (k * G::BY + j) * G::BXP + i;
		#pragma unroll
		for (int q = 0; q < nI; ++q) Us[q * G::BOX + dst] = __ldg(sp.Uin + src + q * g.strideV);
	}
	__syncthreads();

	double const dt = *sp.dt;
	real acc[G::CPT][nI];
	#pragma unroll
	for (int n = 0; n < G::CPT; ++n)
		#pragma unroll
		for (int q = 0; q < nI; ++q) acc[n][q] = 0;

	if (sp.computeL) {
		stageSide<Eqn, DIM, PLM, FLIM, T, 0>(g, sp, ep, Us, SG, FX, acc, real(dt));
		if (DIM >= 2) stageSide<Eqn, DIM, PLM, FLIM, T, 1>(g, sp, ep, Us, SG, FX, acc, real(dt));
		if (DIM >= 3) stageSide<Eqn, DIM, PLM, FLIM, T, 2>(g, sp, ep, Us, SG, FX, acc, real(dt));
	}

	// ---- E: epilogue
	real dtCell = inf_of<real>::v();
	#pragma unroll
	for (int n = 0; n < G::CPT; ++n) {
		int const cell = tid + n * T::NT;
		if (G::CELLS % T::NT != 0 && cell >= G::CELLS) break;
		int const i = cell % G::TX, j = (cell / G::TX) % G::TY, k = cell / (G::TX * G::TY);
		int const gi = i0 + i, gj = j0 + j, gk = k0 + k;
		bool const inside = gi < g.S[0] - HB_G && (DIM < 2 || gj < g.S[1] - HB_G) && (DIM < 3 || gk < g.S[2] - HB_G);
		if (!inside) continue;
		long long const idx = gi + g.strideY * gj + g.strideZ * gk;
		if constexpr (Eqn::eqnId <= 1) {
			// op:addSource of the self-gravity op (solverbase.lua:3219-3223, selfgrav.cl:11-76): accel = central difference of the potential,
			// deriv.m -= accel rho, deriv.ETotal -= m . accel -- applied to L where the reference applies it to derivBuf
			if (sp.gravPot && sp.computeL) {
				real accel[3] = {0, 0, 0};
				accel[0] = (sp.gravPot[idx + 1] - sp.gravPot[idx - 1]) / (real(2.) * g.dx[0]);
				if (DIM >= 2) accel[1] = (sp.gravPot[idx + g.strideY] - sp.gravPot[idx - g.strideY]) / (real(2.) * g.dx[1]);
				if (DIM >= 3) accel[2] = (sp.gravPot[idx + g.strideZ] - sp.gravPot[idx - g.strideZ]) / (real(2.) * g.dx[2]);
				int const box = boxIdx<DIM, T>(i, j, k);
				real const rho = Us[box], mx = Us[G::BOX + box], my = Us[2 * G::BOX + box], mz = Us[3 * G::BOX + box];
				acc[n][1] = acc[n][1] - accel[0] * rho;
				acc[n][2] = acc[n][2] - accel[1] * rho;
				acc[n][3] = acc[n][3] - accel[2] * rho;
				acc[n][4] -= mx * accel[0] + my * accel[1] + mz * accel[2];
			}
		}
		if (sp.Lout) {
			#pragma unroll
			for (int q = 0; q < nI; ++q) sp.Lout[idx + q * g.strideV] = acc[n][q];
		}
		if (!sp.Uout) continue;   // hb_fv_calc_deriv: only dU/dt is wanted
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
			if (sp.computeL) r = r + acc[n][q] * real(sp.betaSelf * dt);
			U[q] = r;
		}
		Eqn::constrainU(ep, U);
		#pragma unroll
		for (int q = 0; q < nI; ++q) sp.Uout[idx + q * g.strideV] = U[q];
		if (sp.dtMinBits) dtCell = rmin<real>(dtCell, Eqn::calcDTCell(ep, U, g.dx, g.dim));
	}
	if (sp.dtMinBits) {
		double v = double(dtCell);
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) { double const u = __shfl_xor_sync(0xffffffffu, v, o); v = u < v ? u : v; }
		if ((tid & 31) == 0) redBuf[tid >> 5] = v;
		__syncthreads();
		if (tid < 32) {
			v = tid < (T::NT + 31) / 32 ? redBuf[tid] : HUGE_VAL;
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) { double const u = __shfl_xor_sync(0xffffffffu, v, o); v = u < v ? u : v; }
			if (tid == 0 && v < HUGE_VAL) atomicMin(sp.dtMinBits, dtBits(v));
		}
	}
}

// ---------------------------------------------------------------------------------------------------
// Ghost fill: boundary_x, then _y, then _z of the reference (gridsolver.lua:1272-1314) compose to a
// per-axis source-index map, so one launch over the ghost cells gives the same result, corners included:
//   periodic  j < g: g + (j - g + 2N) % N ;  j >= S-g (j' = S-1-j): g + (g-1-j') % N   (gridsolver.lua:645,648)
//   mirror    j < g: 2g-1-j ; j >= S-g: 2(S-g)-1-j, negating the reflected vector components (:662-671,741)
//   freeflow  j < g: g ; j >= S-g: S-g-1                                                  (:772-778)
//   none      ghost left untouched (:618-621) -- also used for slab faces owned by a neighbouring rank
//   linear / quadratic / fixed (:746-764, :782-846) do not compose to a source map (a corner ghost is an extrapolation of
//   extrapolated values): a grid with one of those runs the reference's own sequence, fill_ghosts_axis for x, then y, then z.
struct BcP { int bc[6]; const double* fixedState; /* device, [6][HB_FIXED_STRIDE]: what a 'fixed' face writes */ };
constexpr int HB_FIXED_STRIDE = 64;
#ifndef HB_BC_PERIODIC   // same values as include/hydrob200.h
#define HB_BC_PERIODIC 0
#define HB_BC_MIRROR 1
#define HB_BC_FREEFLOW 2
#define HB_BC_NONE 3
#define HB_BC_LINEAR 4
#define HB_BC_QUADRATIC 5
#define HB_BC_FIXED 6
#endif

HB_HD int ghostSource(int j, int S, int bcMin, int bcMax, bool& flip, bool& skip) {
	int const g = HB_G, N = S - 2 * g;
	flip = false; skip = false;
	if (j < g) {
		switch (bcMin) {
		case HB_BC_PERIODIC: return g + (j - g + 2 * N) % N;
		case HB_BC_MIRROR: flip = true; return 2 * g - 1 - j;
		case HB_BC_FREEFLOW: return g;
		default: skip = true; return j;
		}
	} else if (j >= S - g) {
		int const jp = S - 1 - j;
		switch (bcMax) {
		case HB_BC_PERIODIC: return g + (g - 1 - jp) % N;
		case HB_BC_MIRROR: flip = true; return 2 * (S - g) - 1 - j;
		case HB_BC_FREEFLOW: return S - g - 1;
		default: skip = true; return j;
		}
	}
	return j;
}

// Enumerates the ghost cells: z slabs (whole planes), then y slabs of the remaining planes, then x slabs.
template<class Eqn, int MODE>
__global__ void fill_ghosts(GridP<typename Eqn::real> const g, BcP const bc, typename Eqn::real* __restrict__ U, int nVars, int skipRimAxis)
{
	typedef typename Eqn::real real;
	long long const S0 = g.S[0], S1 = g.S[1], S2 = g.S[2];
	int const gy = g.dim >= 2 ? HB_G : 0, gz = g.dim >= 3 ? HB_G : 0;
	long long const nZ = 2LL * gz * S0 * S1;
	long long const nY = 2LL * gy * S0 * (S2 - 2 * gz);
	long long const nX = 2LL * HB_G * (S1 - 2 * gy) * (S2 - 2 * gz);
	long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (w >= nZ + nY + nX) return;
	int i, j, k;
	if (w < nZ) {
		i = int(w % S0); j = int((w / S0) % S1); int kk = int(w / (S0 * S1));
		k = kk < gz ? kk : int(S2) - 2 * gz + kk;
	} else if (w < nZ + nY) {
		w -= nZ;
		i = int(w % S0); int jj = int((w / S0) % (2 * gy)); k = gz + int(w / (S0 * 2 * gy));
		j = jj < gy ? jj : int(S1) - 2 * gy + jj;
	} else {
		w -= nZ + nY;
		int ii = int(w % (2 * HB_G)); j = gy + int((w / (2 * HB_G)) % (S1 - 2 * gy)); k = gz + int(w / ((2 * HB_G) * (S1 - 2 * gy)));
		i = ii < HB_G ? ii : int(S0) - 2 * HB_G + ii;
	}
	// An axis whose method is 'none' leaves the cell as it is, as the reference's pass for that axis would
	// (later passes would still copy it from an interior row; a cell that is ghost along a 'none' axis and along
	// another axis is owned by the neighbouring rank's exchange in the slab-decomposed case, so leave it too).
	bool fx = false, fy = false, fz = false, sx = false, sy = false, sz = false;
	int const si = ghostSource(i, g.S[0], bc.bc[0], bc.bc[1], fx, sx);
	int const sj = g.dim >= 2 ? ghostSource(j, g.S[1], bc.bc[2], bc.bc[3], fy, sy) : j;
	int const sk = g.dim >= 3 ? ghostSource(k, g.S[2], bc.bc[4], bc.bc[5], fz, sz) : k;
	if (sx || sy || sz) return;
	if (skipRimAxis >= 0) {
		// overlapped slab exchange: the ghost cells of the planes that are sent to the neighbours were filled by fill_ghosts_planes
		int const c = skipRimAxis == 1 ? j : k, S = g.S[skipRimAxis];
		if ((c >= HB_G && c < 2 * HB_G) || (c >= S - 2 * HB_G && c < S - HB_G)) return;
	}
	long long const dst = i + g.strideY * j + g.strideZ * k;
	long long const src = si + g.strideY * sj + g.strideZ * sk;
	for (int q = 0; q < nVars; ++q) {
		real v = U[src + q * g.strideV];
		if (fx && Eqn::mirrorFlips(q, 0)) v = real(-1.) * v;
		if (fy && Eqn::mirrorFlips(q, 1)) v = real(-1.) * v;
		if (fz && Eqn::mirrorFlips(q, 2)) v = real(-1.) * v;
		U[dst + q * g.strideV] = v;
	}
}

// The ghost cells (along the axes other than `axis`) of the 2 x numGhost planes next to the slab faces along `axis`: the planes the
// slab exchange sends.  Same source map as fill_ghosts; one thread per cell of those planes.
template<class Eqn, int MODE>
__global__ void fill_ghosts_planes(GridP<typename Eqn::real> const g, BcP const bc, typename Eqn::real* __restrict__ U, int nVars, int axis)
{
	typedef typename Eqn::real real;
	long long const S0 = g.S[0], Sm = axis == 2 ? g.S[1] : 1;          // extent of a plane: S0 x S1 (axis z) or S0 (axis y)
	long long const w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (w >= 2LL * HB_G * S0 * Sm) return;
	int const i = int(w % S0), m = int((w / S0) % Sm), pp = int(w / (S0 * Sm));
	int const S = g.S[axis];
	int const c = pp < HB_G ? HB_G + pp : S - 2 * HB_G + (pp - HB_G);
	int const j = axis == 2 ? m : c, k = axis == 2 ? c : 0;
	bool fx = false, fy = false, fz = false, sx = false, sy = false, sz = false;
	int const si = ghostSource(i, g.S[0], bc.bc[0], bc.bc[1], fx, sx);
	int const sj = (g.dim >= 2 && axis != 1) ? ghostSource(j, g.S[1], bc.bc[2], bc.bc[3], fy, sy) : j;
	if (sx || sy || sz || (si == i && sj == j)) return;
	long long const dst = i + g.strideY * j + g.strideZ * k;
	long long const src = si + g.strideY * sj + g.strideZ * k;
	for (int q = 0; q < nVars; ++q) {
		real v = U[src + q * g.strideV];
		if (fx && Eqn::mirrorFlips(q, 0)) v = real(-1.) * v;
		if (fy && Eqn::mirrorFlips(q, 1)) v = real(-1.) * v;
		U[dst + q * g.strideV] = v;
	}
}

// One pass of the reference's boundary kernel for one axis (boundary_x / _y / _z, gridsolver.lua:1100-1190): one thread per cell
// of the perpendicular plane, ghost cells of the other axes included, `for j < numGhost { min face; max face }` in that order, so
// that the extrapolating methods see the ghost values written one iteration earlier exactly as the reference's loop does.
template<class Eqn, int MODE>
__global__ void fill_ghosts_axis(GridP<typename Eqn::real> const g, BcP const bc, typename Eqn::real* __restrict__ U, int nVars, int axis)
{
	typedef typename Eqn::real real;
	int const o1 = axis == 0 ? 1 : 0, o2 = axis == 2 ? 1 : 2;
	long long const S1 = g.S[o1], S2 = g.S[o2];
	long long const w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (w >= S1 * S2) return;
	long long const strd[3] = {1, g.strideY, g.strideZ};
	long long const base = (w % S1) * strd[o1] + (w / S1) * strd[o2], sa = strd[axis];
	int const S = g.S[axis], G = HB_G, N = S - 2 * G;
	for (int q = 0; q < nVars; ++q) {
		real* const u = U + q * g.strideV + base;
		bool const flips = Eqn::mirrorFlips(q, axis);
		for (int j = 0; j < G; ++j) {
			#pragma unroll
			for (int mm = 0; mm < 2; ++mm) {
				int const method = bc.bc[2 * axis + mm];
				switch (method) {
				case HB_BC_PERIODIC:
					if (mm == 0) u[j * sa] = u[(G + (j - G + 2 * N) % N) * sa];
					else u[(S - 1 - j) * sa] = u[(G + (G - 1 - j) % N) * sa];
					break;
				case HB_BC_MIRROR: {
					real v = mm == 0 ? u[(2 * G - 1 - j) * sa] : u[(S - G - 1 - j) * sa];
					if (flips) v = real(-1.) * v;
					u[(mm == 0 ? j : S - G + j) * sa] = v;
					break; }
				case HB_BC_FREEFLOW:
					if (mm == 0) u[j * sa] = u[G * sa];
					else u[(S - G + j) * sa] = u[(S - G - 1) * sa];
					break;
				case HB_BC_LINEAR:
					if (mm == 0) u[(G - j - 1) * sa] = real(2.) * u[(G - j) * sa] - u[(G - j + 1) * sa];
					else u[(S - G + j) * sa] = real(2.) * u[(S - G + j - 1) * sa] - u[(S - G + j - 2) * sa];
					break;
				case HB_BC_QUADRATIC:
					if (mm == 0) u[(G - j - 1) * sa] = real(3.) * u[(G - j) * sa] - real(3.) * u[(G - j + 1) * sa] + u[(G - j + 2) * sa];
					else u[(S - G + j) * sa] = real(3.) * u[(S - G + j - 1) * sa] - real(3.) * u[(S - G + j - 2) * sa] + u[(S - G + j - 3) * sa];
					break;
				case HB_BC_FIXED:
					u[(mm == 0 ? j : S - G + j) * sa] = real(bc.fixedState[(2 * axis + mm) * HB_FIXED_STRIDE + q]);
					break;
				default: break;
				}
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------------
// Stand-alone kernels: calcDT (eqn.lua:1187-1224 + reduceMin, solverbase.lua:1350-1358, 3004-3023),
// constrainU on every cell (solverbase.lua:2116-2127), AoS <-> SoA conversion at the API boundary.
template<class Eqn, int MODE>
__global__ void calc_dt(GridP<typename Eqn::real> const g, typename Eqn::Params const ep,
	typename Eqn::real const* __restrict__ U, unsigned long long* dtMinBits)
{
	typedef typename Eqn::real real;
	__shared__ double redBuf[32];
	long long const nInt = (long long)g.N[0] * g.N[1] * g.N[2];
	real dtCell = inf_of<real>::v();
	for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < nInt; w += (long long)gridDim.x * blockDim.x) {
		int const i = int(w % g.N[0]), j = int((w / g.N[0]) % g.N[1]), k = int(w / ((long long)g.N[0] * g.N[1]));
		long long const idx = (i + HB_G) + g.strideY * (j + (g.dim >= 2 ? HB_G : 0)) + g.strideZ * (k + (g.dim >= 3 ? HB_G : 0));
		real u[Eqn::nI];
		#pragma unroll
		for (int q = 0; q < Eqn::nI; ++q) u[q] = U[idx + q * g.strideV];
		dtCell = rmin<real>(dtCell, Eqn::calcDTCell(ep, u, g.dx, g.dim));
	}
	double v = double(dtCell);
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) { double const u = __shfl_xor_sync(0xffffffffu, v, o); v = u < v ? u : v; }
	if ((threadIdx.x & 31) == 0) redBuf[threadIdx.x >> 5] = v;
	__syncthreads();
	if (threadIdx.x < 32) {
		v = threadIdx.x < (blockDim.x + 31) / 32 ? redBuf[threadIdx.x] : HUGE_VAL;
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) { double const u = __shfl_xor_sync(0xffffffffu, v, o); v = u < v ? u : v; }
		if (threadIdx.x == 0 && v < HUGE_VAL) atomicMin(dtMinBits, dtBits(v));
	}
}

template<class Eqn, int MODE>
__global__ void constrain_all(GridP<typename Eqn::real> const g, typename Eqn::Params const ep, typename Eqn::real* __restrict__ U)
{
	typedef typename Eqn::real real;
	long long const n = (long long)g.S[0] * g.S[1] * g.S[2];
	long long const w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (w >= n) return;
	int const i = int(w % g.S[0]), j = int((w / g.S[0]) % g.S[1]), k = int(w / ((long long)g.S[0] * g.S[1]));
	long long const idx = i + g.strideY * j + g.strideZ * k;
	real u[Eqn::nI];
	#pragma unroll
	for (int q = 0; q < Eqn::nI; ++q) u[q] = U[idx + q * g.strideV];
	Eqn::constrainU(ep, u);
	#pragma unroll
	for (int q = 0; q < Eqn::nI; ++q) U[idx + q * g.strideV] = u[q];
}

// AoS double[cell][nS] (the reference's cons_t order, INDEX order) <-> SoA real
template<class real>
__global__ void aos_to_soa(GridP<real> const g, int nS, const double* __restrict__ aos, real* __restrict__ U)
{
	long long const n = (long long)g.S[0] * g.S[1] * g.S[2];
	long long const w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (w >= n * nS) return;
	long long const cell = w / nS; int const q = int(w - cell * nS);
	int const i = int(cell % g.S[0]), j = int((cell / g.S[0]) % g.S[1]), k = int(cell / ((long long)g.S[0] * g.S[1]));
	U[i + g.strideY * j + g.strideZ * k + q * g.strideV] = real(aos[w]);
}
template<class real>
__global__ void soa_to_aos(GridP<real> const g, int nS, const real* __restrict__ U, double* __restrict__ aos)
{
	long long const n = (long long)g.S[0] * g.S[1] * g.S[2];
	long long const w = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (w >= n * nS) return;
	long long const cell = w / nS; int const q = int(w - cell * nS);
	int const i = int(cell % g.S[0]), j = int((cell / g.S[0]) % g.S[1]), k = int(cell / ((long long)g.S[0] * g.S[1]));
	aos[w] = double(U[i + g.strideY * j + g.strideZ * k + q * g.strideV]);
}

}   // namespace hb
