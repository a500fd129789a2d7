// hb_fv_ops.h -- table of launchers one (equation, real, fp-mode) translation unit exports to the host API.
// The kernels are compiled once per table (hb_fv_inst.cu with -DHB_EQN/-DHB_REAL/-DHB_STRICT):
//   fast   : nvcc default floating point (FMA contraction on) -- the production kernels
//   strict : -fmad=false -- same source, no contraction; bit-comparable with the non-contracted CPU oracle
#pragma once
#include <cuda_runtime.h>
#include "hb_fv_kernels.cuh"
#include "hb_fv_march.cuh"
#include "hb_fv_march2d.cuh"
#include "hb_fv_march3.cuh"
#include "hb_ops_kernels.cuh"
#include "hb_ctu_kernels.cuh"

namespace hb {

// kernels launched by the last FvOps::stage call on this thread (a stage of the ADM equation is several launches); read by hb_fv.cu
// for hb_fv_launch_count
extern thread_local int tlsStageLaunches;

constexpr int kMarchGenBase = 1000;

template<class real> struct FvOps {
	int eqnId, nS, nI, nW;
	cudaError_t (*stage)(int dim, bool plm, bool flim, GridP<real> const& g, StageP<real> const& sp, const double* eqnParams, cudaStream_t st);
	// Plane-marching TMA kernel (hb_fv_march.cuh).  marchInfo: is it built for (dim, plm, flim, slope limiter)?  If so it
	// returns the TMA box {x, y, z, var} the host must encode into the tensor map of every stage-input buffer and
	// info = {TX, TY, planes per CTA, threads, dynamic smem bytes without staged RK operands, column threads, epilogue adds the self-gravity source}.
	// cfg selects a tile configuration (0 = default).
	bool (*marchInfo)(int dim, bool plm, bool flim, int slopeLimiter, int cfg, int box[4], int info[7]);
	// the GENERAL marching configurations (March3Cfg::GEN: any slope limiter, no reconstruction, Roe with a flux limiter, HLL / Rusanov / HLLC;
	// 3-D): cfg = kMarchGenBase + i, launched through `march` like the others.  Null when not built (ADM, run-time equations).
	bool (*marchInfoGen)(int dim, int cfg, int box[4], int info[7]);
	// chunkSel: 0 all chunks along the marching axis, 1 first + last chunk, 2 the chunks in between (overlapped slab exchange)
	cudaError_t (*march)(int dim, int slopeLimiter, int cfg, const CUtensorMap* tmap, int padX, GridP<real> const& g, StageP<real> const& sp,
		const double* eqnParams, int chunkSel, cudaStream_t st);
	// rimAxis < 0: every ghost cell.  rimAxis = 1 | 2 with planesOnly: only the ghost cells of the planes the slab exchange sends;
	// with !planesOnly: every ghost cell but those.  rimAxis = -2 - axis: the reference's pass for that axis alone (fill_ghosts_axis).
	cudaError_t (*ghosts)(GridP<real> const& g, BcP const& bc, real* U, int nVars, int rimAxis, bool planesOnly, cudaStream_t st);
	cudaError_t (*calcDT)(GridP<real> const& g, const double* eqnParams, const real* U, unsigned long long* dtMinBits, cudaStream_t st);
	cudaError_t (*constrainAll)(GridP<real> const& g, const double* eqnParams, real* U, cudaStream_t st);
	void (*tileInfo)(int dim, bool plm, bool flim, int out[5]);   // TX, TY, TZ, NT, dynamic smem bytes
	// unit-test hook: evaluates one device function per item on the GPU (device pointers; doubles in and out)
	cudaError_t (*debugEval)(int kind, int side, int n, const double* eqnParams, const double* aux, const double* in, double* out, cudaStream_t st);
	// optional (null when unused): device scratch the stage needs (StageP::scratch), in reals; the equation's initDerivs kernel
	long long (*scratchElems)(GridP<real> const& g);
	cudaError_t (*initDerivs)(GridP<real> const& g, real* U, cudaStream_t st);
	// optional (null for equations without ops): the kernels of hydro/op (hb_ops_kernels.cuh), which = HB_OPK_*
	cudaError_t (*opKernel)(int which, GridP<real> const& g, OpP<real> const& o, cudaStream_t st);
	// optional: the unfused kernels of the corner-transport-upwind variant (hb_ctu_kernels.cuh), which = HB_CTUK_*
	cudaError_t (*ctuKernel)(int which, GridP<real> const& g, StageP<real> const& sp, CtuP<real> const& c, const double* eqnParams, cudaStream_t st);
};

// exported by hb_fv_inst.cu instantiations
const FvOps<double>* ops_euler_f64_fast();
const FvOps<double>* ops_euler_f64_strict();
const FvOps<float>* ops_euler_f32_fast();
const FvOps<float>* ops_euler_f32_strict();
const FvOps<double>* ops_mhd_f64_fast();
const FvOps<double>* ops_mhd_f64_strict();
const FvOps<float>* ops_mhd_f32_fast();
const FvOps<float>* ops_mhd_f32_strict();
// exported by hb_adm_inst.cu instantiations
const FvOps<double>* ops_adm3d_f64_fast();
const FvOps<double>* ops_adm3d_f64_strict();
const FvOps<float>* ops_adm3d_f32_fast();
const FvOps<float>* ops_adm3d_f32_strict();

}   // namespace hb
