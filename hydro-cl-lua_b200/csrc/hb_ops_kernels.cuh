// hb_ops_kernels.cuh -- the operators either side of the finite-volume step:
//   Jacobi Poisson relaxation   hydro/op/relaxation.lua:152-196, hydro/op/poisson.cl, hydro/op/poisson_jacobi.cl
//   self-gravity                hydro/op/selfgrav.lua, selfgrav.cl (potential = ePot; euler.lua:179-188, mhd.lua:120-122)
//   NoDiv (Jacobi parent)       hydro/op/nodiv.lua (potential = psi, vector = B; mhd.lua:113-119 with noDivPoissonSolver=jacobi)
//
// B200 design: the reference reads the residual back to the host after every Jacobi sweep (relaxation.lua:177) to decide
// whether to stop.  Here the decision stays on the device: op_finish_iter sums the per-block partial residuals in a fixed
// order, writes residual / iteration into OpCtl and raises OpCtl::done; every later sweep of the same relax() returns at
// once.  A relax() is therefore maxIters x (sweep, copy, ghost fill, finish) launches with no host round trip, and the
// whole update stays one graph-capturable stream sequence.  The state is SoA, so a sweep reads one variable (7 points)
// and writes one: 16 B per cell in double, HBM-bound.
#pragma once
#include "hb_fv_kernels.cuh"

namespace hb {

struct OpCtl { int done; int lastIter; double lastResidual; double maxVal; };

constexpr int HB_OP_NT = 256;

template<class real> struct OpP {
	int kind;               // 1 self-gravity, 2 NoDiv
	real* U;                // variable 0 of the state the op works on
	real* writeBuf;         // one variable, same strides
	double* partial;        // per-block partial sums / maxima
	OpCtl* ctl;
	int pot, vec;           // variable index of the potential; first component of the vector field (NoDiv)
	double param;           // self-gravity: gravitationalConstant / unit_m3_per_kg_s2
	double stopEpsilon; int stopOnEpsilon;
	int iter;               // 1-based sweep number (op_finish_iter)
	double volumeWithoutBorder;
	int nBlocks;            // blocks of the all-cells launches (= entries of `partial`)
};

// all-cells enumeration (SETBOUNDS(0,0)): i fastest
template<class real> HB_D bool opCell(GridP<real> const& g, long long w, int& i, int& j, int& k, long long& idx) {
	long long const S0 = g.S[0], S1 = g.S[1], S2 = g.S[2];
	if (w >= S0 * S1 * S2) return false;
	i = int(w % S0); j = int((w / S0) % S1); k = int(w / (S0 * S1));
	idx = i + g.strideY * j + g.strideZ * k;
	return true;
}
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
		source = source + (v[idx + st] - v[idx - st]) * real(.5 / double(g.dx[s]));
	}
	return source;
}

// relax() start: a fresh stop flag
template<class real> __global__ void op_begin(OpCtl* ctl) { ctl->done = 0; ctl->lastIter = 0; }

// poisson.cl:36-53 initPotential: potential = -source on the interior
template<class real> __global__ void op_init_potential(GridP<real> const g, OpP<real> const o) {
	int i, j, k; long long idx;
	if (!opCell(g, blockIdx.x * (long long)blockDim.x + threadIdx.x, i, j, k, idx)) return;
	if (opOOB(g, i, j, k, HB_G, HB_G)) return;
	o.U[o.pot * g.strideV + idx] = -opSource(g, o, i, j, k, idx);
}

template<class real, bool MAX> HB_D void opBlockReduce(double v, double* out) {
	__shared__ double red[HB_OP_NT / 32];
	#pragma unroll
	for (int m = 16; m > 0; m >>= 1) { double const u = __shfl_xor_sync(0xffffffffu, v, m); v = MAX ? (u > v ? u : v) : v + u; }
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
	__syncthreads();
	if (threadIdx.x < 32) {
		v = threadIdx.x < HB_OP_NT / 32 ? red[threadIdx.x] : (MAX ? -HUGE_VAL : 0.);
		#pragma unroll
		for (int m = 16; m > 0; m >>= 1) { double const u = __shfl_xor_sync(0xffffffffu, v, m); v = MAX ? (u > v ? u : v) : v + u; }
		if (threadIdx.x == 0) *out = v;
	}
}

// poisson_jacobi.cl:42-167 solveJacobi on a cartesian grid: cell_dx_j = grid_dx_j and cell->volume = prod grid_dx, so
// volume_intL = volume_intR = .5 (volume + volume).  Ghost cells copy the potential through; residual^2 goes to the block sum.
template<class real> __global__ void op_solve_jacobi(GridP<real> const g, OpP<real> const o) {
	if (o.ctl->done) return;
	int i, j, k; long long idx;
	double res2 = 0;
	if (opCell(g, blockIdx.x * (long long)blockDim.x + threadIdx.x, i, j, k, idx)) {
		real const* pot = o.U + o.pot * g.strideV;
		if (opOOB(g, i, j, k, HB_G, HB_G)) o.writeBuf[idx] = pot[idx];
		else {
			real volume = 1;
			for (int s = 0; s < g.dim; ++s) volume = volume * g.dx[s];
			real const volL = real(.5) * (volume + volume), volR = real(.5) * (volume + volume), volAtX = volume;
			real skewSum = 0;
			for (int s = 0; s < g.dim; ++s) {
				real const dx = g.dx[s];
				long long const st = opStride(g, s);
				skewSum = skewSum + (pot[idx + st] * (volR / (dx * dx)) + pot[idx - st] * (volL / (dx * dx)));   // real_add3 = a + (b + c), math.cl:221
			}
			skewSum = skewSum * (real(1.) / volAtX);
			real diag = 0;
			for (int s = 0; s < g.dim; ++s) { real const dx = g.dx[s]; diag = diag - (volR + volL) / (dx * dx); }
			diag = diag / volAtX;
			real const source = opSource(g, o, i, j, k, idx);
			real const oldU = pot[idx];
			o.writeBuf[idx] = (source - skewSum) * (real(1.) / diag);
			real const residual = (source - skewSum) - diag * oldU;
			res2 = double(residual * residual);
		}
	}
	if (o.stopOnEpsilon) opBlockReduce<real, false>(res2, o.partial + blockIdx.x);
}

// poisson.cl:55-64 copyWriteToPotentialNoGhost
template<class real> __global__ void op_copy_write(GridP<real> const g, OpP<real> const o) {
	if (o.ctl->done) return;
	int i, j, k; long long idx;
	if (!opCell(g, blockIdx.x * (long long)blockDim.x + threadIdx.x, i, j, k, idx)) return;
	if (opOOB(g, i, j, k, HB_G, HB_G)) return;
	o.U[o.pot * g.strideV + idx] = o.writeBuf[idx];
}

// relaxation.lua:176-194: residual = sqrt(reduceSum / volumeWithoutBorder); stop when |residual| <= stopEpsilon.  One block, fixed order.
template<class real> __global__ void op_finish_iter(OpP<real> const o) {
	if (o.ctl->done) return;
	double v = 0;
	if (o.stopOnEpsilon) for (int n = threadIdx.x; n < o.nBlocks; n += HB_OP_NT) v += o.partial[n];
	__shared__ double total;
	opBlockReduce<real, false>(v, &total);
	__syncthreads();
	if (threadIdx.x == 0) {
		o.ctl->lastIter = o.iter;
		if (o.stopOnEpsilon) {
			double const residual = sqrt(double(real(total)) / o.volumeWithoutBorder);
			o.ctl->lastResidual = residual;
			if (fabs(residual) <= o.stopEpsilon) o.ctl->done = 1;
		}
	}
}

// selfgrav.lua:123-147 offsetPotential: the potential minus its maximum over ALL cells (copyPotentialToReduce is SETBOUNDS(0,0))
template<class real> __global__ void op_max_partial(GridP<real> const g, OpP<real> const o) {
	int i, j, k; long long idx;
	double v = -HUGE_VAL;
	if (opCell(g, blockIdx.x * (long long)blockDim.x + threadIdx.x, i, j, k, idx)) v = double(o.U[o.pot * g.strideV + idx]);
	opBlockReduce<real, true>(v, o.partial + blockIdx.x);
}
template<class real> __global__ void op_max_finish(OpP<real> const o) {
	double v = -HUGE_VAL;
	for (int n = threadIdx.x; n < o.nBlocks; n += HB_OP_NT) { double const u = o.partial[n]; v = u > v ? u : v; }
	opBlockReduce<real, true>(v, &o.ctl->maxVal);
}
template<class real> __global__ void op_offset(GridP<real> const g, OpP<real> const o) {
	int i, j, k; long long idx;
	if (!opCell(g, blockIdx.x * (long long)blockDim.x + threadIdx.x, i, j, k, idx)) return;
	real* p = o.U + o.pot * g.strideV + idx;
	*p = *p - real(o.ctl->maxVal);
}

// nodiv.lua:133-157 noDiv: B -= grad psi (central differences) on the interior
template<class real> __global__ void op_nodiv(GridP<real> const g, OpP<real> const o) {
	int i, j, k; long long idx;
	if (!opCell(g, blockIdx.x * (long long)blockDim.x + threadIdx.x, i, j, k, idx)) return;
	if (opOOB(g, i, j, k, HB_G, HB_G)) return;
	real const* pot = o.U + o.pot * g.strideV;
	for (int s = 0; s < g.dim; ++s) {
		long long const st = opStride(g, s);
		real const dv = (pot[idx + st] - pot[idx - st]) * real(1. / (2. * double(g.dx[s])));
		real* v = o.U + (o.vec + s) * g.strideV + idx;
		*v = *v - dv;
	}
}

enum { HB_OPK_BEGIN = 0, HB_OPK_INIT, HB_OPK_JACOBI, HB_OPK_COPY, HB_OPK_FINISH, HB_OPK_MAX_PARTIAL, HB_OPK_MAX_FINISH, HB_OPK_OFFSET, HB_OPK_NODIV };

template<class real> cudaError_t launchOpKernel(int which, GridP<real> const& g, OpP<real> const& o, cudaStream_t st) {
	unsigned const nb = (unsigned)o.nBlocks;
	switch (which) {
	case HB_OPK_BEGIN: op_begin<real><<<1, 1, 0, st>>>(o.ctl); break;
	case HB_OPK_INIT: op_init_potential<real><<<nb, HB_OP_NT, 0, st>>>(g, o); break;
	case HB_OPK_JACOBI: op_solve_jacobi<real><<<nb, HB_OP_NT, 0, st>>>(g, o); break;
	case HB_OPK_COPY: op_copy_write<real><<<nb, HB_OP_NT, 0, st>>>(g, o); break;
	case HB_OPK_FINISH: op_finish_iter<real><<<1, HB_OP_NT, 0, st>>>(o); break;
	case HB_OPK_MAX_PARTIAL: op_max_partial<real><<<nb, HB_OP_NT, 0, st>>>(g, o); break;
	case HB_OPK_MAX_FINISH: op_max_finish<real><<<1, HB_OP_NT, 0, st>>>(o); break;
	case HB_OPK_OFFSET: op_offset<real><<<nb, HB_OP_NT, 0, st>>>(g, o); break;
	case HB_OPK_NODIV: op_nodiv<real><<<nb, HB_OP_NT, 0, st>>>(g, o); break;
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}

}   // namespace hb
