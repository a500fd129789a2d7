// hb_ops_kernels.cuh -- the operators either side of the finite-volume step:
//   Jacobi Poisson relaxation   hydro/op/relaxation.lua:152-196, hydro/op/poisson.cl, hydro/op/poisson_jacobi.cl
//   self-gravity                hydro/op/selfgrav.lua, selfgrav.cl (potential = ePot; euler.lua:179-188, mhd.lua:120-122)
//   NoDiv (Jacobi parent)       hydro/op/nodiv.lua (potential = psi, vector = B; mhd.lua:113-119 with noDivPoissonSolver=jacobi)
//
// B200 design: the reference reads the residual back to the host after every Jacobi sweep (relaxation.lua:177) to decide
// whether to stop.  Here the decision stays on the device: the sweep kernel sums the per-block partial residuals in a fixed
// order, writes residual / iteration into OpCtl and raises OpCtl::done; every later sweep of the same relax() returns at
// once.  A relax() is therefore maxIters x (sweep, ghost fill) launches with no host round trip, and the whole update stays
// one graph-capturable stream sequence.  The residual sum is finished by the sweep's last block (ticket counter, partials
// combined in index order: deterministic).  The state is SoA, so a sweep reads the potential (7 points, one stream from
// HBM / L2), the source operand (rho, or the neighbours of B) and writes the potential: 3-4 words per cell, HBM-bound.
#pragma once
#include "hb_fv_kernels.cuh"

namespace hb {

struct OpCtl { int done; int lastIter; double lastResidual; double maxVal; unsigned int ticket; unsigned int pad; double sumLocal; };

constexpr int HB_OP_NT = 256;

template<class real> struct OpP {
	int kind;               // 1 self-gravity, 2 NoDiv
	real* U;                // variable 0 of the state the op works on
	real* writeBuf;         // one variable, same strides (relaxation.lua:52-57)
	const real* potIn;      // the sweep reads this copy of the potential ...
	real* potOut;           // ... and writes that one (U's potential slot and writeBuf alternate, see relax() in hb_fv.cu)
	double* partial;        // per-block partial sums / maxima
	OpCtl* ctl;
	int pot, vec;           // variable index of the potential; first component of the vector field (NoDiv)
	double param;           // self-gravity: gravitationalConstant / unit_m3_per_kg_s2
	double stopEpsilon; int stopOnEpsilon;
	int iter;               // 1-based sweep number
	double volumeWithoutBorder;
	int nBlocks;            // blocks of the row-strided launches (= entries of `partial`)
	int ctaRows;            // 1: a row per CTA, 0: a row per warp
	int deferDecision;      // slab-decomposed grid: the sweep leaves its residual sum in OpCtl::sumLocal; after the all-reduce op_decide finishes the iteration
	// grid constants of solveJacobi, formed on the host in `real` with the reference's operations (IEEE division: the same bits as on the device)
	real cS[3];             // volume_int / (dx_s dx_s)
	real invVol;            // 1. / volAtX
	real invDiag;           // 1. / diag
	real diag;
	real sDiv[3];           // NoDiv source: .5 / grid_dx_s
	real sGrad[3];          // noDiv kernel: 1. / (2. grid_dx_s)
};

// Rows (j, k) of the ghost-inclusive array are dealt round-robin to WARPS (OpP::ctaRows = 0: many short rows, the 3-D case -- a 260- or
// 516-cell row wastes at most one 32-cell step instead of most of a CTA-wide one) or to CTAs (ctaRows = 1: few long rows, the 2-D case).
// Either way every access is a coalesced segment and the index arithmetic is one integer division per row.
constexpr int HB_OP_WARPS = HB_OP_NT / 32;
#define HB_OP_ROWS(g, row, j, k, base) \
	for (int row = o.ctaRows ? blockIdx.x : blockIdx.x * HB_OP_WARPS + (threadIdx.x >> 5); row < (g).S[1] * (g).S[2]; row += o.ctaRows ? gridDim.x : gridDim.x * HB_OP_WARPS) \
		if (int const j = row % (g).S[1], k = row / (g).S[1]; true) \
			if (long long const base = (g).strideY * j + (g).strideZ * k; true)
#define HB_OP_LANES(g, i) for (int i = o.ctaRows ? threadIdx.x : (threadIdx.x & 31); i < (g).S[0]; i += o.ctaRows ? HB_OP_NT : 32)

template<class real> HB_D bool opOOB(GridP<real> const& g, int i, int j, int k, int l, int r) {
	return i < l || i >= g.S[0] - r || (g.dim >= 2 && (j < l || j >= g.S[1] - r)) || (g.dim >= 3 && (k < l || k >= g.S[2] - r));
}
template<class real> HB_D long long opStride(GridP<real> const& g, int s) { return s == 0 ? 1 : (s == 1 ? g.strideY : g.strideZ); }

// getPoissonDivCode: selfgrav.lua:43-48 (4 pi rho G / unit) | nodiv.lua:86-118 (central divergence of the vector field, 0 next to the rim)
template<class real> HB_D real opSource(GridP<real> const& g, OpP<real> const& o, int i, int j, int k, long long idx) {
	if (o.kind == 1) return real(4. * 3.14159265358979323846 * double(o.U[idx]) * o.param / 1.);
	real source = 0;
	if (opOOB(g, i, j, k, 1, 1)) return source;
	for (int s = 0; s < g.dim; ++s) {
		real const* v = o.U + (o.vec + s) * g.strideV;
		long long const st = opStride(g, s);
		source = source + (v[idx + st] - v[idx - st]) * o.sDiv[s];
	}
	return source;
}

// relax() start: a fresh stop flag
template<class real, int MODE> __global__ void op_begin(OpCtl* ctl) { ctl->done = 0; ctl->lastIter = 0; ctl->ticket = 0; }

// poisson.cl:36-53 initPotential: potential = -source on the interior
template<class real, int MODE> __global__ void op_init_potential(GridP<real> const g, OpP<real> const o) {
	HB_OP_ROWS(g, row, j, k, base)
		HB_OP_LANES(g, i) {
			if (opOOB(g, i, j, k, HB_G, HB_G)) continue;
			o.U[o.pot * g.strideV + base + i] = -opSource(g, o, i, j, k, base + i);
		}
}

template<bool MAX> HB_D double opCombine(double a, double b) { return MAX ? (b > a ? b : a) : a + b; }
// block-wide reduction in a fixed order; the result is valid in thread 0
template<bool MAX> HB_D double opBlockReduce(double v) {
	__shared__ double red[HB_OP_NT / 32];
	#pragma unroll
	for (int m = 16; m > 0; m >>= 1) v = opCombine<MAX>(v, __shfl_xor_sync(0xffffffffu, v, m));
	__syncthreads();
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
	__syncthreads();
	if (threadIdx.x < 32) {
		v = threadIdx.x < HB_OP_NT / 32 ? red[threadIdx.x] : (MAX ? -HUGE_VAL : 0.);
		#pragma unroll
		for (int m = 16; m > 0; m >>= 1) v = opCombine<MAX>(v, __shfl_xor_sync(0xffffffffu, v, m));
	}
	return v;
}
// The block that finishes last (ticket counter) combines the per-block partials in index order: deterministic whichever block it is.
template<bool MAX> HB_D bool opLastBlockReduce(double mine, double* partial, OpCtl* ctl, int nBlocks, double& total) {
	__shared__ int isLast;
	if (threadIdx.x == 0) {
		partial[blockIdx.x] = mine;
		__threadfence();
		unsigned int const t = atomicAdd(&ctl->ticket, 1u);
		isLast = t == (unsigned int)(nBlocks - 1);
	}
	__syncthreads();
	if (!isLast) return false;
	__threadfence();
	double v = MAX ? -HUGE_VAL : 0.;
	for (int n = threadIdx.x; n < nBlocks; n += HB_OP_NT) v = opCombine<MAX>(v, ((volatile double*)partial)[n]);
	total = opBlockReduce<MAX>(v);
	if (threadIdx.x == 0) ctl->ticket = 0;
	return true;
}

// One Jacobi sweep = solveJacobi (poisson_jacobi.cl:42-167, cartesian: cell_dx_j = grid_dx_j, cell->volume = prod grid_dx, so
// volume_intL = volume_intR = .5 (volume + volume)) + copyWriteToPotentialNoGhost (poisson.cl:55-64) + the residual test of
// relaxation.lua:176-194.  The reference writes the sweep to writeBuf (ghost cells: the old potential) and copies its interior back;
// here the two copies of the potential alternate instead (potIn -> potOut; the ghost fill that follows makes potOut what the reference's
// potential is after its boundary pass), which saves the copy kernel: 3 words per cell and sweep instead of 5.
template<class real, int MODE> __global__ void __launch_bounds__(HB_OP_NT) op_solve_jacobi(GridP<real> const g, OpP<real> const o) {
	if (o.ctl->done) return;
	real const* __restrict__ pot = o.potIn;
	real* __restrict__ out = o.potOut;
	double res2 = 0;
	HB_OP_ROWS(g, row, j, k, base)
		HB_OP_LANES(g, i) {
			long long const idx = base + i;
			if (opOOB(g, i, j, k, HB_G, HB_G)) { out[idx] = pot[idx]; continue; }
			real skewSum = 0;
			skewSum = skewSum + (pot[idx + 1] * o.cS[0] + pot[idx - 1] * o.cS[0]);                                         // real_add3 = a + (b + c), math.cl:221
			if (g.dim >= 2) skewSum = skewSum + (pot[idx + g.strideY] * o.cS[1] + pot[idx - g.strideY] * o.cS[1]);
			if (g.dim >= 3) skewSum = skewSum + (pot[idx + g.strideZ] * o.cS[2] + pot[idx - g.strideZ] * o.cS[2]);
			skewSum = skewSum * o.invVol;
			real const source = opSource(g, o, i, j, k, idx);
			real const oldU = pot[idx];
			out[idx] = (source - skewSum) * o.invDiag;
			real const residual = (source - skewSum) - o.diag * oldU;
			res2 += double(residual * residual);
		}
	double const mine = o.stopOnEpsilon ? opBlockReduce<false>(res2) : 0.;
	double total = 0;
	if (!opLastBlockReduce<false>(mine, o.partial, o.ctl, o.nBlocks, total)) return;
	if (threadIdx.x == 0) {
		if (o.deferDecision) { o.ctl->sumLocal = total; return; }
		o.ctl->lastIter = o.iter;
		if (o.stopOnEpsilon) {
			double const residual = sqrt(double(real(total)) / o.volumeWithoutBorder);
			o.ctl->lastResidual = residual;
			if (fabs(residual) <= o.stopEpsilon) o.ctl->done = 1;
		}
	}
}

// Slab-decomposed grid: the iteration's bookkeeping after the residual sums of the slabs have been all-reduced into OpCtl::sumLocal.
// Every rank sees the same sum, so every rank stops at the same sweep.
template<class real, int MODE> __global__ void op_decide(OpP<real> const o) {
	if (o.ctl->done) return;
	o.ctl->lastIter = o.iter;
	if (o.stopOnEpsilon) {
		double const residual = sqrt(double(real(o.ctl->sumLocal)) / o.volumeWithoutBorder);
		o.ctl->lastResidual = residual;
		if (fabs(residual) <= o.stopEpsilon) o.ctl->done = 1;
	}
}

// After the last sweep: an odd number of sweeps leaves the result in writeBuf; bring it home (all cells: its ghost cells are the filled ones)
template<class real, int MODE> __global__ void op_final_copy(GridP<real> const g, OpP<real> const o) {
	if ((o.ctl->lastIter & 1) == 0) return;
	real* __restrict__ pot = o.U + o.pot * g.strideV;
	HB_OP_ROWS(g, row, j, k, base)
		HB_OP_LANES(g, i) pot[base + i] = o.writeBuf[base + i];
}

// selfgrav.lua:123-147 offsetPotential: the potential minus its maximum over ALL cells (copyPotentialToReduce is SETBOUNDS(0,0))
template<class real, int MODE> __global__ void __launch_bounds__(HB_OP_NT) op_max(GridP<real> const g, OpP<real> const o) {
	real const* __restrict__ pot = o.U + o.pot * g.strideV;
	double v = -HUGE_VAL;
	HB_OP_ROWS(g, row, j, k, base)
		HB_OP_LANES(g, i) { double const u = double(pot[base + i]); v = u > v ? u : v; }
	double const mine = opBlockReduce<true>(v);
	double total = 0;
	if (!opLastBlockReduce<true>(mine, o.partial, o.ctl, o.nBlocks, total)) return;
	if (threadIdx.x == 0) o.ctl->maxVal = total;
}
template<class real, int MODE> __global__ void op_offset(GridP<real> const g, OpP<real> const o) {
	real* __restrict__ pot = o.U + o.pot * g.strideV;
	real const m = real(o.ctl->maxVal);
	HB_OP_ROWS(g, row, j, k, base)
		HB_OP_LANES(g, i) pot[base + i] = pot[base + i] - m;
}

// nodiv.lua:133-157 noDiv: B -= grad psi (central differences) on the interior
template<class real, int MODE> __global__ void op_nodiv(GridP<real> const g, OpP<real> const o) {
	real const* __restrict__ pot = o.U + o.pot * g.strideV;
	HB_OP_ROWS(g, row, j, k, base)
		HB_OP_LANES(g, i) {
			if (opOOB(g, i, j, k, HB_G, HB_G)) continue;
			long long const idx = base + i;
			for (int s = 0; s < g.dim; ++s) {
				long long const st = opStride(g, s);
				real const dv = (pot[idx + st] - pot[idx - st]) * o.sGrad[s];
				real* v = o.U + (o.vec + s) * g.strideV + idx;
				*v = *v - dv;
			}
		}
}

enum { HB_OPK_BEGIN = 0, HB_OPK_INIT, HB_OPK_JACOBI, HB_OPK_FINAL_COPY, HB_OPK_MAX, HB_OPK_OFFSET, HB_OPK_NODIV, HB_OPK_DECIDE };

// MODE (0 production, 1 strict = -fmad=false) only makes the kernels of the two builds distinct symbols: without it the linker would
// merge the equally named instantiations of the two translation units and one build would run the other's code.
template<class real, int MODE> cudaError_t launchOpKernel(int which, GridP<real> const& g, OpP<real> const& o, cudaStream_t st) {
	unsigned const nb = (unsigned)o.nBlocks;
	switch (which) {
	case HB_OPK_BEGIN: op_begin<real, MODE><<<1, 1, 0, st>>>(o.ctl); break;
	case HB_OPK_INIT: op_init_potential<real, MODE><<<nb, HB_OP_NT, 0, st>>>(g, o); break;
	case HB_OPK_JACOBI: op_solve_jacobi<real, MODE><<<nb, HB_OP_NT, 0, st>>>(g, o); break;
	case HB_OPK_FINAL_COPY: op_final_copy<real, MODE><<<nb, HB_OP_NT, 0, st>>>(g, o); break;
	case HB_OPK_MAX: op_max<real, MODE><<<nb, HB_OP_NT, 0, st>>>(g, o); break;
	case HB_OPK_OFFSET: op_offset<real, MODE><<<nb, HB_OP_NT, 0, st>>>(g, o); break;
	case HB_OPK_NODIV: op_nodiv<real, MODE><<<nb, HB_OP_NT, 0, st>>>(g, o); break;
	case HB_OPK_DECIDE: op_decide<real, MODE><<<1, 1, 0, st>>>(o); break;
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}

}   // namespace hb
