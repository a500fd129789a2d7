// hb_adm_inst.cu -- instantiates the ADM Bona-Masso 3-D finite-volume kernels for one (real, fp-mode).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -DHB_REAL=double -DHB_OPS=ops_adm3d_f64_fast [-fmad=false -DHB_STRICT=1]
#include "hb_fv_ops.h"
#include "hb_adm_kernels.cuh"
#include <cstdlib>

namespace hb {
namespace {

typedef HB_REAL real;
#ifdef HB_STRICT
constexpr int MODE = 1;
#else
constexpr int MODE = 0;
#endif
#ifdef HB_STRICT
typedef ADM3D<real, false> Eqn;     // literal arithmetic, -fmad=false: bit-comparable with the oracle
#else
typedef ADM3D<real, true> Eqn;      // production forms (reciprocal / rsqrt seeds, ratio-free minmod / superbee)
#endif

template<int SIDE, int V>
cudaError_t launchFluxShared(GridP<real> const& g, StageP<real> const& sp, Eqn::Params const& ep, cudaStream_t st) {
	typedef AdmFluxGeom<SIDE, V> G;
	int const nOut = G::NB - 2;
	dim3 grid;
	if (SIDE == 0) grid = dim3((unsigned)((g.N[0] + 1 + nOut - 1) / nOut), (unsigned)g.N[1], (unsigned)g.N[2]);
	else if (SIDE == 1) grid = dim3((unsigned)((g.N[0] + G::LX - 1) / G::LX), (unsigned)((g.N[1] + 1 + nOut - 1) / nOut), (unsigned)g.N[2]);
	else grid = dim3((unsigned)((g.N[0] + G::LX - 1) / G::LX), (unsigned)((g.N[2] + 1 + nOut - 1) / nOut), (unsigned)g.N[1]);
	adm_flux_shared<Eqn, SIDE, MODE, V><<<grid, G::NT, 0, st>>>(g, ep, sp.Uin, sp.scratch, sp.dt, sp.fluxLimiter);
	return cudaGetLastError();
}

template<int SIDE>
cudaError_t launchFlux(GridP<real> const& g, StageP<real> const& sp, Eqn::Params const& ep, cudaStream_t st) {
	static int const shared = getenv("HB_ADM_SHARED") ? atoi(getenv("HB_ADM_SHARED")) : 3;   // measured at 128^3: 0 2.42, 1 2.18, 2 2.30, 3 2.15 ms per stage
	if (shared && sp.fluxLimiter > 0 && g.fluxOn[SIDE]) {
		if (shared == 2) return launchFluxShared<SIDE, 2>(g, sp, ep, st);
		if (shared == 3) return launchFluxShared<SIDE, 3>(g, sp, ep, st);
		if (shared == 5) return launchFluxShared<SIDE, 5>(g, sp, ep, st);
		return launchFluxShared<SIDE, 1>(g, sp, ep, st);
	}
	long long const n = (long long)(g.N[0] + (SIDE == 0)) * (g.N[1] + (SIDE == 1)) * (g.N[2] + (SIDE == 2));
	int const nt = 128;
	adm_flux<Eqn, SIDE, MODE><<<(unsigned)((n + nt - 1) / nt), nt, 0, st>>>(g, ep, sp.Uin, sp.scratch, sp.dt, sp.fluxLimiter);
	return cudaGetLastError();
}

// plm / flim are ignored: the equation runs the Roe flux on cell-centred states with the flux limiter given in sp (the reference's
// configuration for this equation; hb_fv_create rejects usePLM).  Needs sp.scratch = 3 x 13 x strideV reals.
cudaError_t stage(int dim, bool, bool, GridP<real> const& g, StageP<real> const& sp, const double* eqnParams, cudaStream_t st) {
	Eqn::Params const ep = Eqn::makeParams(eqnParams);
	cudaError_t e = cudaSuccess;
	if (sp.computeL) {
		if (!sp.scratch) return cudaErrorInvalidValue;
		e = launchFlux<0>(g, sp, ep, st);
		if (e == cudaSuccess && dim >= 2) e = launchFlux<1>(g, sp, ep, st);
		if (e == cudaSuccess && dim >= 3) e = launchFlux<2>(g, sp, ep, st);
		if (e != cudaSuccess) return e;
	}
	long long const n = (long long)g.N[0] * g.N[1] * g.N[2];
	int const nt = 128;
	static int const split = getenv("HB_ADM_SPLIT") ? atoi(getenv("HB_ADM_SPLIT")) : 1;
	tlsStageLaunches = (sp.computeL ? dim : 0) + (split == 2 ? 4 : (split ? 6 : 1));
	unsigned const nb = (unsigned)((n + nt - 1) / nt);
	if (split == 2) {
		adm_update<Eqn, MODE, 0><<<nb, nt, 0, st>>>(g, sp, ep, sp.scratch);
		adm_update<Eqn, MODE, 7><<<nb, nt, 0, st>>>(g, sp, ep, sp.scratch);
		adm_update<Eqn, MODE, 1><<<nb, nt, 0, st>>>(g, sp, ep, sp.scratch);
		adm_update<Eqn, MODE, 3><<<nb, nt, 0, st>>>(g, sp, ep, sp.scratch);
	} else if (split) {
		adm_update<Eqn, MODE, 0><<<nb, nt, 0, st>>>(g, sp, ep, sp.scratch);
		adm_update<Eqn, MODE, 7><<<nb, nt, 0, st>>>(g, sp, ep, sp.scratch);
		adm_update<Eqn, MODE, 1><<<nb, nt, 0, st>>>(g, sp, ep, sp.scratch);
		adm_update<Eqn, MODE, 4><<<nb, nt, 0, st>>>(g, sp, ep, sp.scratch);
		adm_update<Eqn, MODE, 5><<<nb, nt, 0, st>>>(g, sp, ep, sp.scratch);
		adm_update<Eqn, MODE, 6><<<nb, nt, 0, st>>>(g, sp, ep, sp.scratch);
	} else {
		adm_update<Eqn, MODE, 2><<<(unsigned)((n + nt - 1) / nt), nt, 0, st>>>(g, sp, ep, sp.scratch);
	}
	return cudaGetLastError();
}
bool marchInfo(int, bool, bool, int, int, int*, int*) { return false; }
cudaError_t march(int, int, int, const CUtensorMap*, int, GridP<real> const&, StageP<real> const&, const double*, int, cudaStream_t) { return cudaErrorInvalidValue; }

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
	int const nt = 128;
	long long blocks = (n + nt - 1) / nt;
	if (blocks > 148 * 16) blocks = 148 * 16;
	calc_dt<Eqn, MODE><<<(unsigned)blocks, nt, 0, st>>>(g, Eqn::makeParams(ep), U, dtMinBits);
	return cudaGetLastError();
}
cudaError_t constrainAll(GridP<real> const&, const double*, real*, cudaStream_t) { return cudaSuccess; }   // diagnostics only: nothing to do
void tileInfo(int, bool, bool, int out[5]) { out[0] = 128; out[1] = 1; out[2] = 1; out[3] = 128; out[4] = 0; }
cudaError_t debugEval(int, int, int, const double*, const double*, const double*, double*, cudaStream_t) { return cudaErrorInvalidValue; }
long long scratchElems(GridP<real> const& g) { return 3LL * 13 * g.strideV; }
cudaError_t initDerivs(GridP<real> const& g, real* U, cudaStream_t st) {
	long long const n = (long long)g.N[0] * g.N[1] * g.N[2];
	int const nt = 128;
	adm_init_derivs<Eqn, MODE><<<(unsigned)((n + nt - 1) / nt), nt, 0, st>>>(g, U);
	return cudaGetLastError();
}

const FvOps<real> theOps = {Eqn::eqnId, Eqn::nS, Eqn::nI, Eqn::nW, stage, marchInfo, nullptr, march, ghosts, calcDT, constrainAll, tileInfo, debugEval,
	scratchElems, initDerivs};

}   // namespace

const FvOps<real>* HB_OPS() { return &theOps; }

}   // namespace hb
