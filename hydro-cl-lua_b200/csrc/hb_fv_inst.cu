// hb_fv_inst.cu -- instantiates the fused stage kernel family for one (equation, real, fp-mode).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -DHB_EQN=Euler -DHB_REAL=double -DHB_OPS=ops_euler_f64_fast [-fmad=false]
#include "hb_fv_ops.h"

#ifndef HB_EQN
#error "compile with -DHB_EQN=<Euler|MHD> -DHB_REAL=<double|float> -DHB_OPS=<exported name>"
#endif

namespace hb {
namespace {

typedef HB_REAL real;
#ifdef HB_STRICT
constexpr int MODE = 1;
typedef HB_EQN<real, false> Eqn;     // literal arithmetic, -fmad=false: bit-comparable with the oracle
#else
constexpr int MODE = 0;
typedef HB_EQN<real, true> Eqn;      // production arithmetic (hb_roe_fast.cuh) in the marching kernel
#endif

// Tile shapes (interior cells per CTA) and CTA size per dimensionality.
typedef Tile<256, 1, 1, 256> Tile1;
typedef Tile<32, 16, 1, 256> Tile2;
typedef Tile<16, 8, 4, 256> Tile3;
template<int DIM> struct TileFor;
template<> struct TileFor<1> { typedef Tile1 type; };
template<> struct TileFor<2> { typedef Tile2 type; };
template<> struct TileFor<3> { typedef Tile3 type; };

template<int DIM, bool PLM, bool FLIM>
cudaError_t launchStage(GridP<real> const& g, StageP<real> const& sp, const double* eqnParams, cudaStream_t st) {
	typedef typename TileFor<DIM>::type T;
	typedef TileGeom<DIM, T> G;
	auto kern = fv_stage<Eqn, DIM, PLM, FLIM, T, MODE>;
	size_t const smem = G::template smemBytes<real, Eqn::nI>(PLM ? (sp.plmMode >= 2 ? 2 : 1) : 0);
	if (sp.plmMode >= 2 && !Eqn::hasEigenForCell) return cudaErrorInvalidValue;
	constexpr size_t kLimit = 232448 - 1024;
	if (smem > kLimit) return cudaErrorInvalidConfiguration;     // 'plm athena' keeps both face states: 3-D MHD does not fit this tile
	static bool attrSet = false;
	if (!attrSet) {
		size_t const athena = G::template smemBytes<real, Eqn::nI>(PLM ? 2 : 0), cons = G::template smemBytes<real, Eqn::nI>(PLM ? 1 : 0);
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(Eqn::hasEigenForCell && athena <= kLimit ? athena : cons));
		if (e != cudaSuccess) return e;
		attrSet = true;
	}
	long long const ntx = (g.N[0] + G::TX - 1) / G::TX;
	long long const nty = (g.N[1] + G::TY - 1) / G::TY;
	long long const ntz = (g.N[2] + G::TZ - 1) / G::TZ;
	long long const blocks = ntx * nty * ntz;
	kern<<<(unsigned)blocks, T::NT, smem, st>>>(g, sp, Eqn::makeParams(eqnParams));
	return cudaGetLastError();
}

template<int DIM>
cudaError_t launchStageDim(bool plm, bool flim, GridP<real> const& g, StageP<real> const& sp, const double* ep, cudaStream_t st) {
	if (plm) return launchStage<DIM, true, false>(g, sp, ep, st);
	if (flim) return launchStage<DIM, false, true>(g, sp, ep, st);
	return launchStage<DIM, false, false>(g, sp, ep, st);
}

cudaError_t stage(int dim, bool plm, bool flim, GridP<real> const& g, StageP<real> const& sp, const double* ep, cudaStream_t st) {
	switch (dim) {
	case 1: return launchStageDim<1>(plm, flim, g, sp, ep, st);
	case 2: return launchStageDim<2>(plm, flim, g, sp, ep, st);
	case 3: return launchStageDim<3>(plm, flim, g, sp, ep, st);
	}
	return cudaErrorInvalidValue;
}

// ---- plane-marching kernel: tile configurations per dimensionality (cfg index; the host takes the first one whose shared
// memory fits, starting at $HB_MARCH_CFG or 0).  X(index, WX, TY, KM, MINB, VAR)
#ifdef HB_STRICT
// the strict build carries fewer configurations (compile time)
#define HB_MARCH3_LIST(X) X(0, 1, 8, 64, 1, 16) X(1, 1, 8, 32, 1, 0) X(2, 1, 4, 32, 1, 0) X(3, 1, 8, 64, 1, 48)
#define HB_MARCH2_LIST(X) X(0, 4, 1, 32, 2, 0) X(1, 3, 1, 32, 2, 0)
#else
// Measured on B200, 512 x 512 x 128 Euler double RK4 stage (profiles/r01c_sweep_c4.txt): cfg 0 2.74 ms, cfg 1 2.80, cfg 4 2.74, cfg 5 3.46,
// cfg 6 (three flux cores issued as one block) 4.13: longer operand live ranges cost more than the interleaving gains.
#define HB_MARCH3_LIST(X) \
	X(0, 1, 8, 64, 1, 16)   /* 32 x 8 columns, 64 planes per CTA, 11 warps, staggered slope phase */ \
	X(1, 1, 8, 32, 1, 0)    /* 32 x 8 columns, 32 planes per CTA */ \
	X(2, 1, 4, 32, 1, 0)    /* 32 x 4 columns, 7 warps: the fallback that fits 8-variable equations (MHD) in shared memory */ \
	X(3, 1, 10, 32, 1, 0)   /* 32 x 10 columns, 13 warps (fits with at most two staged RK operands) */ \
	X(4, 1, 8, 128, 1, 16)  /* 128 planes per CTA */ \
	X(5, 1, 4, 32, 2, 0)    /* 32 x 4 columns, 7 warps, 2 CTAs / SM */ \
	X(6, 1, 8, 32, 1, 1)    /* column warps issue their three flux cores as one block (kept for the record: slower) */ \
	X(7, 1, 8, 64, 1, 48)   /* cfg 0 + the self-gravity source in the epilogue (MarchCfg::GRAV; chosen by hb_fv_add_op, never by the auto selection) */
// 2-D, 2048^2 stage: Euler cfg 0 0.266 ms, cfg 1 0.28; MHD cfg 0 0.730 ms (168 registers, spills), cfg 1 0.706 (254 registers, none)
#define HB_MARCH2_LIST(X) \
	X(0, 4, 1, 32, 2, 0)    /* 128 columns, 5 warps, 2 CTAs / SM */ \
	X(1, 3, 1, 32, 2, 0)    /* 96 columns, 4 warps */ \
	X(2, 2, 1, 32, 4, 0)    /* 64 columns, 3 warps */ \
	X(3, 6, 1, 64, 1, 0)    /* 192 columns, 7 warps (TMA boxes are at most 256 elements wide) */
#endif
// 3-D second-generation marching kernel (hb_fv_march3.cuh): X(index, TY, KM, VAR); these come first in the 3-D cfg numbering
#ifdef HB_STRICT
#define HB_MARCH3N_LIST(X) X(0, 15, 64, 0) X(1, 11, 64, 0) X(2, 8, 32, 0) X(3, 6, 64, 0) X(4, 15, 64, 32) X(5, 11, 64, 32) X(6, 8, 32, 32) X(7, 6, 64, 32)
#else
#if defined(HB_DEV)
#define HB_MARCH3N_LIST(X) X(0, 15, 64, 0) X(1, 11, 64, 0) X(2, 15, 64, 16) X(3, 11, 64, 16)
#else
#define HB_MARCH3N_LIST(X) \
	/* VAR 16 (March3Cfg::OPTMA): the RK operands of the epilogue are staged by TMA -- one bulk copy per operand and plane, issued by the halo warp \
	   a plane ahead behind an 'operand area free' barrier -- instead of per-thread cp.async: measured 2.18 -> 1.99 ms per 512^3 stage (r02m) */ \
	X(0, 15, 64, 16)   /* 32 x 15 columns, 64 planes per CTA, 16 warps at 128 registers (fits with at most two staged RK operands) */ \
	X(1, 11, 64, 16)   /* 32 x 11 columns, 12 warps at 168 registers (classic RK4's four operands fit) */ \
	X(2, 7, 64, 16)    /* 32 x 7 columns, 8 warps = 256 threads: an 8-variable equation (ideal MHD, ~226 registers) keeps all its registers -- 32 x 8 is \
	                      9 warps, allocated as 384 threads, i.e. capped at 168 registers with spills: measured 28 % slower per row (profiles/r02p_sweep_mhd_ty7.txt) */ \
	X(3, 6, 64, 16)    /* 32 x 6 columns, 7 warps */ \
	X(4, 15, 128, 16)  /* 128 planes per CTA */ \
	X(5, 11, 32, 16)   /* 32 planes per CTA */ \
	X(6, 15, 64, 32)   /* the self-gravity source in the epilogue, cp.async operands (chosen by hb_fv_add_op, never by the auto selection) */ \
	X(7, 11, 64, 32) \
	X(8, 7, 64, 32) \
	X(9, 6, 64, 32) \
	X(10, 15, 64, 0)   /* the round-2 baseline: operands staged by per-thread cp.async ($HB_MARCH_CFG=10 runs 10,10,10,11 for an A/B comparison) */ \
	X(11, 11, 64, 0) \
	X(12, 15, 64, 8)   /* operands read from global memory in the epilogue (no staging): measured slower, kept for the record */ \
	X(13, 8, 64, 16)   /* 32 x 8 columns, 9 warps (round 2's MHD tile until call p; kept for the comparison) */
	/* PAIR (VAR bit 0: the x and y cores of a cell as one block) with TMA operands, X(14, 7, 64, 17) X(15, 11, 64, 17) X(16, 15, 64, 17): 2.30 / 2.05 / 2.16 ms
	   against 2.43 / 2.02 / 1.94 without it -- it pays only where registers are free (8 warps), and 16 warps win (profiles/r02w_sweep_pair.txt) */
	/* two CTAs per SM (March3Cfg::MINB; X(13, 7, 64, 80), X(14, 6, 64, 80)) measured slower: 2.03 / 2.18 against 1.93 ms (profiles/r02n_sweep_minb2.txt) */
#endif
#endif
// general configurations (March3Cfg::GEN, VAR bit 1): X(index, TY, KM, VAR); cfg = kMarchGenBase + index
#define HB_MARCH3G_LIST(X) X(0, 8, 64, 2) X(1, 6, 64, 2) X(2, 4, 32, 2)   /* (32 x 7 measured slower here: profiles/r02q_gen_vs_tile.txt) */
// the same for 2-D (March2Cfg::GEN, MINB bit 5): X(index, NW, KM, MINB)
#define HB_MARCH2G_LIST(X) X(0, 4, 32, 33) X(1, 2, 32, 33)
constexpr int kMarch3N =
#define HB_X(i, ty, km, var) +1
	0 HB_MARCH3N_LIST(HB_X);
#undef HB_X
// 2-D warp-per-pencil kernel (hb_fv_march2d.cuh): X(index, NW, KM, MINB); these come first in the 2-D cfg numbering
#ifdef HB_STRICT
#define HB_MARCH2W_LIST(X) X(0, 4, 32, 1) X(1, 4, 32, 17)
#else
// MINB + 16: cfg 0 with the self-gravity source in the epilogue (March2Cfg::GRAV; chosen by hb_fv_add_op, never by the auto selection)
#define HB_MARCH2W_LIST(X) X(0, 4, 32, 1) X(1, 2, 32, 1) X(2, 8, 32, 1) X(3, 4, 32, 17)
#endif
constexpr int kMarch2W =
#define HB_X(i, nw, km, mb) +1
	0 HB_MARCH2W_LIST(HB_X);
#undef HB_X
// per-equation default: the host starts at cfg 0; 2-D MHD prefers list entry 1
constexpr int remapCfg(int dim, int cfg) {
	if (dim != 2 || cfg < kMarch2W) return cfg;
	int const c = cfg - kMarch2W;                                // index into HB_MARCH2_LIST
	return kMarch2W + ((Eqn::eqnId == 1 && c < 2) ? 1 - c : c);
}

constexpr size_t kSmemLimit = 232448 - 1024;   // 227 KB opt-in maximum per CTA minus the kernel's static shared memory (rounded up)

template<int DIM, int LIM, class C>
cudaError_t launchMarch(const CUtensorMap* tmap, int padX, GridP<real> const& g, StageP<real> const& sp, const double* eqnParams, int chunkSel, cudaStream_t st) {
	typedef MarchGeom<DIM, C, real> G;
	auto kern = fv_march<Eqn, DIM, LIM, C, MODE>;
	int nOps = sp.nOps;                                 // alpha terms on other buffers + beta terms (+ the folded last stage's running sum)
	size_t const smem = G::template smemBytes<Eqn::nI>(nOps);
	size_t const smemMax = G::template smemBytes<Eqn::nI>(2 * HB_MAX_TERMS);
	static bool attrSet = false;
	if (!attrSet) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smemMax < kSmemLimit ? smemMax : kSmemLimit));
		if (e != cudaSuccess) return e;
		attrSet = true;
	}
	if (smem > kSmemLimit) return cudaErrorInvalidConfiguration;
	long long const ntx = (g.N[0] + G::TX - 1) / G::TX;
	long long const nty = DIM == 3 ? (g.N[1] + G::TY - 1) / G::TY : 1;
	long long nm = (g.N[DIM - 1] + C::KM - 1) / C::KM;
	if (chunkSel == 1) nm = nm < 2 ? nm : 2;
	else if (chunkSel == 2) nm = nm > 2 ? nm - 2 : 0;
	if (nm == 0) return cudaSuccess;
	kern<<<(unsigned)(ntx * nty * nm), G::NT, smem, st>>>(*tmap, g, sp, Eqn::makeParams(eqnParams), padX, chunkSel);
	return cudaGetLastError();
}
template<int DIM, class C>
cudaError_t launchMarchLim(int lim, const CUtensorMap* tmap, int padX, GridP<real> const& g, StageP<real> const& sp, const double* ep, int chunkSel, cudaStream_t st) {
	if (lim == 8) return launchMarch<DIM, 8, C>(tmap, padX, g, sp, ep, chunkSel, st);      // minmod
	if (lim == 18) return launchMarch<DIM, 18, C>(tmap, padX, g, sp, ep, chunkSel, st);    // superbee
	return cudaErrorInvalidValue;
}
template<int LIM, class C>
cudaError_t launchMarch3(const CUtensorMap* tmap, int padX, GridP<real> const& g, StageP<real> const& sp, const double* eqnParams, int chunkSel, cudaStream_t st) {
	typedef March3Geom<C, real> G;
	auto kern = fv_march3<Eqn, LIM, C, MODE>;
	int nOps = sp.nOps;                                 // alpha terms on other buffers + beta terms (+ the folded last stage's running sum)
	if (C::OPDIRECT) nOps = 0;                          // operands are read from global memory in the epilogue: no staging area
	size_t const smem = G::template smemBytes<Eqn::nI>(nOps);
	size_t const smemMax = G::template smemBytes<Eqn::nI>(C::OPDIRECT ? 0 : 2 * HB_MAX_TERMS);
	static bool attrSet = false;
	if (!attrSet) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smemMax < kSmemLimit ? smemMax : kSmemLimit));
		if (e != cudaSuccess) return e;
		if (C::MINB > 1) {                                // two CTAs per SM need the whole shared-memory carveout
			e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
			if (e != cudaSuccess) return e;
		}
		attrSet = true;
	}
	if (smem > kSmemLimit) return cudaErrorInvalidConfiguration;
	long long const ntx = (g.N[0] + G::TX - 1) / G::TX;
	long long const nty = (g.N[1] + G::TY - 1) / G::TY;
	long long nm = (g.N[2] + C::KM - 1) / C::KM;
	if (chunkSel) {                                     // rim / interior split of the overlapped slab exchange (see the kernel)
		if (g.N[2] < 2 * HB_G + 1) return cudaErrorInvalidConfiguration;
		nm = chunkSel == 1 ? 2 : (g.N[2] - 2 * HB_G + C::KM - 1) / C::KM;
	}
	kern<<<(unsigned)(ntx * nty * nm), G::NT, smem, st>>>(*tmap, g, sp, Eqn::makeParams(eqnParams), padX, chunkSel);
	return cudaGetLastError();
}
template<class C>
cudaError_t launchMarch3Lim(int lim, const CUtensorMap* tmap, int padX, GridP<real> const& g, StageP<real> const& sp, const double* ep, int chunkSel, cudaStream_t st) {
	if constexpr (C::GEN) return launchMarch3<-1, C>(tmap, padX, g, sp, ep, chunkSel, st);      // limiter, reconstruction and flux at run time
	else {
		if (lim == 8) return launchMarch3<8, C>(tmap, padX, g, sp, ep, chunkSel, st);
		if (lim == 18) return launchMarch3<18, C>(tmap, padX, g, sp, ep, chunkSel, st);
		return cudaErrorInvalidValue;
	}
}
template<class C> void march3InfoCfg(int box[4], int info[7]) {
	typedef March3Geom<C, real> G;
	box[0] = G::BX; box[1] = G::BY; box[2] = 1; box[3] = Eqn::nI;
	info[0] = G::TX; info[1] = G::TY; info[2] = C::KM; info[3] = G::NT; info[4] = (int)G::template smemBytes<Eqn::nI>(0);
	info[5] = C::OPDIRECT ? 0 : G::NCOL * 32; info[6] = (C::GRAV ? 1 : 0) | 2 | 4 | (C::OPTMA ? 8 : 0);   // bit 3: operands by TMA (StageP::opMaps)   // bit 1: fv_march3; bit 2: chunkSel = rim / interior planes
}
template<int LIM, class C>
cudaError_t launchMarch2W(const CUtensorMap* tmap, int padX, GridP<real> const& g, StageP<real> const& sp, const double* eqnParams, int chunkSel, cudaStream_t st) {
	typedef March2Geom<C, real> G;
	auto kern = fv_march2d<Eqn, LIM, C, MODE>;
	int nOps = sp.nOps;                                 // alpha terms on other buffers + beta terms (+ the folded last stage's running sum)
	size_t const smem = G::template smemBytes<Eqn::nI>(nOps);
	size_t const smemMax = G::template smemBytes<Eqn::nI>(2 * HB_MAX_TERMS);
	static bool attrSet = false;
	if (!attrSet) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smemMax < kSmemLimit ? smemMax : kSmemLimit));
		if (e != cudaSuccess) return e;
		attrSet = true;
	}
	if (smem > kSmemLimit) return cudaErrorInvalidConfiguration;
	long long const nSeg = (g.N[0] + G::CW - 1) / G::CW;
	// rows per warp.  The warps are independent, so the grid is nSeg x ceil(N1 / KM) warps; a short chunk costs one extra interface
	// row per KM rows, a long one leaves the last warps of a small grid alone on the GPU.  Measured on 2048^2 (profiles/r01c_sweep_2d_warp.txt):
	// Euler 16 rows 0.253 ms, 32 rows 0.279, 48 rows 0.303; MHD 0.543 / 0.547 / 0.553 (64).  16 rows until the grid fills the GPU
	// eight times over, then 32.  $HB_MARCH2_KM overrides.
	static int kmCache = 0; static long long kmFor[2] = {0, 0};
	if (!kmCache || kmFor[0] != g.N[0] || kmFor[1] != g.N[1]) {
		int perSM = 0, dev = 0, sms = 148;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, G::NT, smem) != cudaSuccess || perSM < 1) perSM = 1;
		long long const conc = (long long)sms * perSM * C::NW;
		int km = nSeg * ((g.N[1] + 15) / 16) <= 8 * conc ? 16 : 32;
		if (const char* e = getenv("HB_MARCH2_KM")) if (atoi(e) >= 2) km = atoi(e);
		kmCache = km; kmFor[0] = g.N[0]; kmFor[1] = g.N[1];
	}
	int const KM = kmCache;
	long long nm = (g.N[1] + KM - 1) / KM;
	if (chunkSel) {                                     // rim / interior split of the overlapped slab exchange (see the kernel)
		if (g.N[1] < 2 * HB_G + 1) return cudaErrorInvalidConfiguration;
		nm = chunkSel == 1 ? 2 : (g.N[1] - 2 * HB_G + KM - 1) / KM;
	}
	long long const blocks = (nSeg * nm + C::NW - 1) / C::NW;
	kern<<<(unsigned)blocks, G::NT, smem, st>>>(*tmap, g, sp, Eqn::makeParams(eqnParams), padX, chunkSel, KM);
	return cudaGetLastError();
}
template<class C>
cudaError_t launchMarch2WLim(int lim, const CUtensorMap* tmap, int padX, GridP<real> const& g, StageP<real> const& sp, const double* ep, int chunkSel, cudaStream_t st) {
	if constexpr (C::GEN) return launchMarch2W<-1, C>(tmap, padX, g, sp, ep, chunkSel, st);      // limiter, reconstruction and flux at run time
	else {
		if (lim == 8) return launchMarch2W<8, C>(tmap, padX, g, sp, ep, chunkSel, st);
		if (lim == 18) return launchMarch2W<18, C>(tmap, padX, g, sp, ep, chunkSel, st);
		return cudaErrorInvalidValue;
	}
}
template<class C> void march2WInfoCfg(int box[4], int info[7]) {
	typedef March2Geom<C, real> G;
	box[0] = G::BX; box[1] = 1; box[2] = 1; box[3] = Eqn::nI;
	info[0] = G::CW * C::NW; info[1] = 1; info[2] = C::KM; info[3] = G::NT; info[4] = (int)G::template smemBytes<Eqn::nI>(0);
	info[5] = C::NW * 32; info[6] = (C::GRAV ? 1 : 0) | 4;   // bit 2: chunkSel = rim / interior rows
}
template<int DIM, class C> void marchInfoCfg(int box[4], int info[7]) {
	typedef MarchGeom<DIM, C, real> G;
	box[0] = G::BX; box[1] = DIM == 3 ? G::BY : 1; box[2] = 1; box[3] = Eqn::nI;
	info[0] = G::TX; info[1] = G::TY; info[2] = C::KM; info[3] = G::NT; info[4] = (int)G::template smemBytes<Eqn::nI>(0);
	info[5] = G::NREG * 32; info[6] = C::GRAV ? 1 : 0;
}
bool marchInfo(int dim, bool plm, bool flim, int lim, int cfg, int box[4], int info[7]) {
	if (!plm || flim || (lim != 8 && lim != 18) || dim < 2 || cfg < 0) return false;
	cfg = remapCfg(dim, cfg);
#define HB_X(i, ty, km, var) if (dim == 3 && cfg == i) { march3InfoCfg<March3Cfg<ty, km, var>>(box, info); return true; }
	HB_MARCH3N_LIST(HB_X)
#undef HB_X
#define HB_X(i, wx, ty, km, mb, var) if (cfg == kMarch3N + i) { marchInfoCfg<3, MarchCfg<wx, ty, km, mb, var>>(box, info); return true; }
	if (dim == 3) { HB_MARCH3_LIST(HB_X) return false; }
#undef HB_X
#define HB_X(i, nw, km, mb) if (cfg == i) { march2WInfoCfg<March2Cfg<nw, km, mb>>(box, info); return true; }
	HB_MARCH2W_LIST(HB_X)
#undef HB_X
#define HB_X(i, wx, ty, km, mb, var) if (cfg == kMarch2W + i) { marchInfoCfg<2, MarchCfg<wx, ty, km, mb, var>>(box, info); return true; }
	HB_MARCH2_LIST(HB_X)
#undef HB_X
	return false;
}
bool marchInfoGen(int dim, int cfg, int box[4], int info[7]) {
#define HB_X(i, nw, km, mb) if (dim == 2 && cfg == kMarchGenBase + i) { march2WInfoCfg<March2Cfg<nw, km, mb>>(box, info); return true; }
	HB_MARCH2G_LIST(HB_X)
#undef HB_X
	if (dim != 3) return false;
#define HB_X(i, ty, km, var) if (cfg == kMarchGenBase + i) { march3InfoCfg<March3Cfg<ty, km, var>>(box, info); return true; }
	HB_MARCH3G_LIST(HB_X)
#undef HB_X
	return false;
}
cudaError_t march(int dim, int lim, int cfg, const CUtensorMap* tmap, int padX, GridP<real> const& g, StageP<real> const& sp, const double* ep, int chunkSel, cudaStream_t st) {
#define HB_X(i, ty, km, var) if (dim == 3 && cfg == kMarchGenBase + i) return launchMarch3Lim<March3Cfg<ty, km, var>>(lim, tmap, padX, g, sp, ep, chunkSel, st);
	HB_MARCH3G_LIST(HB_X)
#undef HB_X
#define HB_X(i, nw, km, mb) if (dim == 2 && cfg == kMarchGenBase + i) return launchMarch2WLim<March2Cfg<nw, km, mb>>(lim, tmap, padX, g, sp, ep, chunkSel, st);
	HB_MARCH2G_LIST(HB_X)
#undef HB_X
	cfg = remapCfg(dim, cfg);
#define HB_X(i, ty, km, var) if (dim == 3 && cfg == i) return launchMarch3Lim<March3Cfg<ty, km, var>>(lim, tmap, padX, g, sp, ep, chunkSel, st);
	HB_MARCH3N_LIST(HB_X)
#undef HB_X
#define HB_X(i, wx, ty, km, mb, var) if (cfg == kMarch3N + i) return launchMarchLim<3, MarchCfg<wx, ty, km, mb, var>>(lim, tmap, padX, g, sp, ep, chunkSel, st);
	if (dim == 3) { HB_MARCH3_LIST(HB_X) }
#undef HB_X
#define HB_X(i, nw, km, mb) if (cfg == i) return launchMarch2WLim<March2Cfg<nw, km, mb>>(lim, tmap, padX, g, sp, ep, chunkSel, st);
	else if (dim == 2) { HB_MARCH2W_LIST(HB_X) }
#undef HB_X
#define HB_X(i, wx, ty, km, mb, var) if (cfg == kMarch2W + i) return launchMarchLim<2, MarchCfg<wx, ty, km, mb, var>>(lim, tmap, padX, g, sp, ep, chunkSel, st);
	if (dim == 2) { HB_MARCH2_LIST(HB_X) }
#undef HB_X
	return cudaErrorInvalidValue;
}

template<int DIM> void tileInfoDim(bool plm, int out[5]) {
	typedef typename TileFor<DIM>::type T;
	typedef TileGeom<DIM, T> G;
	out[0] = G::TX; out[1] = G::TY; out[2] = G::TZ; out[3] = T::NT;
	out[4] = (int)G::template smemBytes<real, Eqn::nI>(plm ? 1 : 0);
}
void tileInfo(int dim, bool plm, bool, int out[5]) {
	if (dim == 1) tileInfoDim<1>(plm, out);
	else if (dim == 2) tileInfoDim<2>(plm, out);
	else tileInfoDim<3>(plm, out);
}

cudaError_t ghosts(GridP<real> const& g, BcP const& bc, real* U, int nVars, int rimAxis, bool planesOnly, cudaStream_t st) {
	long long const S0 = g.S[0], S1 = g.S[1], S2 = g.S[2];
	int const nt = 256;
	if (rimAxis <= -2) {   // one pass of the per-axis sequence (extrapolating / fixed methods): axis = -2 - rimAxis
		int const axis = -2 - rimAxis;
		long long const n = axis == 0 ? S1 * S2 : (axis == 1 ? S0 * S2 : S0 * S1);
		fill_ghosts_axis<Eqn, MODE><<<(unsigned)((n + nt - 1) / nt), nt, 0, st>>>(g, bc, U, nVars, axis);
		return cudaGetLastError();
	}
	if (rimAxis >= 0 && planesOnly) {
		long long const n = 2LL * HB_G * S0 * (rimAxis == 2 ? S1 : 1);
		fill_ghosts_planes<Eqn, MODE><<<(unsigned)((n + nt - 1) / nt), nt, 0, st>>>(g, bc, U, nVars, rimAxis);
		return cudaGetLastError();
	}
	int const gy = g.dim >= 2 ? HB_G : 0, gz = g.dim >= 3 ? HB_G : 0;
	long long const n = 2LL * gz * S0 * S1 + 2LL * gy * S0 * (S2 - 2 * gz) + 2LL * HB_G * (S1 - 2 * gy) * (S2 - 2 * gz);
	fill_ghosts<Eqn, MODE><<<(unsigned)((n + nt - 1) / nt), nt, 0, st>>>(g, bc, U, nVars, rimAxis);
	return cudaGetLastError();
}

cudaError_t calcDT(GridP<real> const& g, const double* ep, const real* U, unsigned long long* dtMinBits, cudaStream_t st) {
	long long const n = (long long)g.N[0] * g.N[1] * g.N[2];
	int const nt = 256;
	long long blocks = (n + nt - 1) / nt;
	if (blocks > 148 * 16) blocks = 148 * 16;
	calc_dt<Eqn, MODE><<<(unsigned)blocks, nt, 0, st>>>(g, Eqn::makeParams(ep), U, dtMinBits);
	return cudaGetLastError();
}

cudaError_t constrainAll(GridP<real> const& g, const double* ep, real* U, cudaStream_t st) {
	long long const n = (long long)g.S[0] * g.S[1] * g.S[2];
	int const nt = 256;
	constrain_all<Eqn, MODE><<<(unsigned)((n + nt - 1) / nt), nt, 0, st>>>(g, Eqn::makeParams(ep), U);
	return cudaGetLastError();
}

// ---- unit-test hook (hb_debug_eval): kind 0 roeFlux(UL,UR) ; 1 constrainU ; 2 calcDTCell (aux = dx[3], dim) ;
//      3 plmHalfSlope (aux[0] = limiter id; in = UL,U,UR) ; 4 roeFluxLimited (aux = limiter id, dt/dx; in = U2L,UL,UR,U2R)
template<int SIDE, int MODE_>
__global__ void debug_eval(int kind, int n, Eqn::Params const ep, const double* __restrict__ aux, const double* __restrict__ in, double* __restrict__ out)
{
	constexpr int nI = Eqn::nI;
	int const w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= n) return;
	if (kind == 0) {
		real UL[nI], UR[nI], F[nI];
		for (int q = 0; q < nI; ++q) { UL[q] = real(in[w * 2 * nI + q]); UR[q] = real(in[w * 2 * nI + nI + q]); }
		roeFlux<Eqn, SIDE>(F, ep, UL, UR);
		for (int q = 0; q < nI; ++q) out[w * nI + q] = double(F[q]);
	} else if (kind == 1) {
		real U[nI];
		for (int q = 0; q < nI; ++q) U[q] = real(in[w * nI + q]);
		Eqn::constrainU(ep, U);
		for (int q = 0; q < nI; ++q) out[w * nI + q] = double(U[q]);
	} else if (kind == 2) {
		real U[nI]; real dx[3] = {real(aux[0]), real(aux[1]), real(aux[2])};
		for (int q = 0; q < nI; ++q) U[q] = real(in[w * nI + q]);
		out[w] = double(Eqn::calcDTCell(ep, U, dx, int(aux[3])));
	} else if (kind == 3) {
		out[w] = double(plmHalfSlope<real>(int(aux[0]), real(in[w * 3]), real(in[w * 3 + 1]), real(in[w * 3 + 2])));
	} else if (kind == 4) {
		real U[4][nI], F[nI];
		for (int c = 0; c < 4; ++c) for (int q = 0; q < nI; ++q) U[c][q] = real(in[(w * 4 + c) * nI + q]);
		roeFluxLimited<Eqn, SIDE>(F, ep, int(aux[0]), real(aux[1]), U[0], U[1], U[2], U[3]);
		for (int q = 0; q < nI; ++q) out[w * nI + q] = double(F[q]);
	}
}
cudaError_t debugEval(int kind, int side, int n, const double* ep, const double* aux, const double* in, double* out, cudaStream_t st) {
	int const nt = 128, nb = (n + nt - 1) / nt;
	if (side == 0) debug_eval<0, MODE><<<nb, nt, 0, st>>>(kind, n, Eqn::makeParams(ep), aux, in, out);
	else if (side == 1) debug_eval<1, MODE><<<nb, nt, 0, st>>>(kind, n, Eqn::makeParams(ep), aux, in, out);
	else debug_eval<2, MODE><<<nb, nt, 0, st>>>(kind, n, Eqn::makeParams(ep), aux, in, out);
	return cudaGetLastError();
}

const FvOps<real> theOps = {Eqn::eqnId, Eqn::nS, Eqn::nI, Eqn::nW, stage, marchInfo, marchInfoGen, march, ghosts, calcDT, constrainAll, tileInfo, debugEval, nullptr, nullptr, launchOpKernel<real, MODE>, launchCtuKernel<Eqn, MODE>};

}   // namespace

const FvOps<real>* HB_OPS() { return &theOps; }

}   // namespace hb
