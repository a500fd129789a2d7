// hb_fv_inst.cu -- instantiates the fused stage kernel family for one (equation, real, fp-mode).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -DHB_EQN=Euler -DHB_REAL=double -DHB_OPS=ops_euler_f64_fast [-fmad=false]
#include "hb_fv_ops.h"

#ifndef HB_EQN
#error "compile with -DHB_EQN=<Euler|MHD> -DHB_REAL=<double|float> -DHB_OPS=<exported name>"
#endif

namespace hb {
namespace {

typedef HB_REAL real;
typedef HB_EQN<real> Eqn;
#ifdef HB_STRICT
constexpr int MODE = 1;
#else
constexpr int MODE = 0;
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
	size_t const smem = G::template smemBytes<real, Eqn::nI>(PLM);
	static bool attrSet = false;
	if (!attrSet) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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

template<int DIM> void tileInfoDim(bool plm, int out[5]) {
	typedef typename TileFor<DIM>::type T;
	typedef TileGeom<DIM, T> G;
	out[0] = G::TX; out[1] = G::TY; out[2] = G::TZ; out[3] = T::NT;
	out[4] = (int)G::template smemBytes<real, Eqn::nI>(plm);
}
void tileInfo(int dim, bool plm, bool, int out[5]) {
	if (dim == 1) tileInfoDim<1>(plm, out);
	else if (dim == 2) tileInfoDim<2>(plm, out);
	else tileInfoDim<3>(plm, out);
}

cudaError_t ghosts(GridP<real> const& g, BcP const& bc, real* U, int nVars, cudaStream_t st) {
	long long const S0 = g.S[0], S1 = g.S[1], S2 = g.S[2];
	int const gy = g.dim >= 2 ? HB_G : 0, gz = g.dim >= 3 ? HB_G : 0;
	long long const n = 2LL * gz * S0 * S1 + 2LL * gy * S0 * (S2 - 2 * gz) + 2LL * HB_G * (S1 - 2 * gy) * (S2 - 2 * gz);
	int const nt = 256;
	fill_ghosts<Eqn, MODE><<<(unsigned)((n + nt - 1) / nt), nt, 0, st>>>(g, bc, U, nVars);
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

const FvOps<real> theOps = {Eqn::eqnId, Eqn::nS, Eqn::nI, Eqn::nW, stage, ghosts, calcDT, constrainAll, tileInfo, debugEval};

}   // namespace

const FvOps<real>* HB_OPS() { return &theOps; }

}   // namespace hb
