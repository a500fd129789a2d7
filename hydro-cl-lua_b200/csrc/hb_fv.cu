// hb_fv.cu -- host side of the fused finite-volume path of the C ABI (hb_fv_* in include/hydrob200.h).
//
// What the reference's LuaJIT host does per solver:update() (hydro/solver/solverbase.lua:3026-3238,
// hydro/int/rk.lua:47-167, hydro/int/fe.lua:33-49) is turned into a static per-stage plan at creation time:
//   stage i (0-based) reads U^i with its ghosts, forms L^i = dU/dt(U^i) on chip, and writes
//       U^{i+1} = sum_{k<=i} alpha[i][k] U^k + dt sum_{k<=i} beta[i][k] L^k      (alpha terms first, k ascending)
//   already constrained (constrainU), followed by one ghost fill; L^i goes to HBM only if a later stage needs it;
//   the last stage also produces the CFL minimum for the next step.  dt and t live on the device
//   (`ctl`), so a whole update is one stream-ordered sequence without host read-back; it can be replayed as a CUDA graph.
// Slab decomposition (one process per GPU): the slowest used axis is split; after each ghost fill the two
// boundary plane pairs are exchanged with ncclSend/ncclRecv and dt is min-allreduced (choppedup.lua:193-233,344-409).
#include "hb_core.h"
#include "hb_fv_ops.h"
#include "hb_jit.h"
#include <cuda.h>
#include <dlfcn.h>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <sstream>

namespace hb {

thread_local int tlsStageLaunches = 1;

// Step bookkeeping on the device (solverbase.lua:3160-3173): dt = fixedDT | cfl * min ; t += dt.
// Runs at the top of every update so that a whole update is one stream-ordered (graph-capturable) sequence
// with no host read-back.  `ctl` = { t, dt, cfl, fixedDT(<0: adaptive) }.
__global__ void begin_step(double* ctl, unsigned long long* dtMinBits)
{
	double dt;
	if (ctl[3] >= 0) dt = ctl[3];
	else dt = ctl[2] * __longlong_as_double((long long)*dtMinBits);
	ctl[1] = dt;
	ctl[0] = ctl[0] + dt;
	*dtMinBits = dtBits(HUGE_VAL);
}


__global__ void set_step(double* ctl, unsigned long long* dtMinBits, double dt) {
	ctl[1] = dt;
	*dtMinBits = dtBits(HUGE_VAL);
}
__global__ void reset_dtmin(unsigned long long* dtMinBits) { *dtMinBits = dtBits(HUGE_VAL); }

// ---- NCCL, bound at run time (the library is the one the host process already loaded, e.g. torch's)
struct Nccl {
	typedef struct ncclComm* comm_t;
	struct uid { char internal[128]; };
	int (*GetUniqueId)(uid*) = nullptr;
	int (*CommInitRank)(comm_t*, int, uid, int) = nullptr;
	int (*CommDestroy)(comm_t) = nullptr;
	int (*Send)(const void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
	int (*Recv)(void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
	int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
	bool ok = false;
	std::string why;
	static Nccl& get() {
		static Nccl n;
		static bool tried = false;
		if (!tried) {
			tried = true;
			void* h = nullptr;
			const char* names[] = {"libnccl.so.2", "libnccl.so"};
			for (const char* nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
			if (!h) { n.why = std::string("cannot load libnccl: ") + dlerror(); return n; }
#define HB_SYM(field, name) *(void**)(&n.field) = dlsym(h, name); if (!n.field) { n.why = std::string("libnccl lacks ") + name; return n; }
			HB_SYM(GetUniqueId, "ncclGetUniqueId") HB_SYM(CommInitRank, "ncclCommInitRank") HB_SYM(CommDestroy, "ncclCommDestroy")
			HB_SYM(Send, "ncclSend") HB_SYM(Recv, "ncclRecv") HB_SYM(AllReduce, "ncclAllReduce")
			HB_SYM(GroupStart, "ncclGroupStart") HB_SYM(GroupEnd, "ncclGroupEnd") HB_SYM(GetErrorString, "ncclGetErrorString")
#undef HB_SYM
			n.ok = true;
		}
		return n;
	}
};
enum { kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2, kNcclMin = 3 };
#define HB_NCCL(expr) do { int r_ = (expr); if (r_ != 0) return setError(HB_ERR_CUDA, std::string(#expr) + ": " + Nccl::get().GetErrorString(r_)); } while (0)

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links neither libcuda nor libnvrtc)
typedef CUresult (*EncodeTiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
	const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled_t encodeTiledFn() {
	static EncodeTiled_t fn = nullptr;
	static bool tried = false;
	if (!tried) {
		tried = true;
		cudaDriverEntryPointQueryResult q;
		void* p = nullptr;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p) fn = (EncodeTiled_t)p;
		else cudaGetLastError();
	}
	return fn;
}

struct Term { int k; double coef; };
struct StagePlan {
	int uIn, uOut;                 // physical U buffers
	int lOut;                      // physical L buffer or -1
	std::vector<Term> alpha;       // k = physical U buffer
	std::vector<Term> beta;        // k = physical L buffer
	double betaSelf; bool computeL; bool last;
	int readsU, readsL;            // words per cell (x nI) for hb_fv_describe
	int marchCfg = -1;             // marching-kernel configuration of this stage (-1: the solver's): a stage with few RK operands fits a larger tile
	bool opTma = false;            // its configuration stages the RK operands by TMA (StageP::opMaps)
	// folded last stage (foldFinalStage): this stage also writes the running sum of the last stage's combination
	int accOut = -1, accIn = -1;   // physical U buffers (accIn < 0 with accOut >= 0: the sum starts here, from the stage input)
	double accCoef = 0, accBetaSelf = 0;
	int operands() const {         // RK operands a marching kernel stages in shared memory
		int n = (int)beta.size() + (accIn >= 0 && accOut >= 0 ? 1 : 0);
		for (auto& t : alpha) if (t.k != uIn) ++n;
		return n;
	}
};

// Static plan of one update for a Butcher tableau (hydro/int/rk.lua:17-44 decides the same "needed later" sets).
static void buildPlan(int order, const double* A, const double* B, std::vector<StagePlan>& plan, int& nU, int& nL) {
	int const n = order < 1 ? 1 : order;
	auto a = [&](int i, int k) { return order < 1 ? 1. : A[i * order + k]; };
	auto b = [&](int i, int k) { return order < 1 ? 1. : B[i * order + k]; };
	std::vector<int> uPhys(n + 1, -1), lPhys(n, -1);
	std::set<int> liveU, liveL;
	uPhys[0] = 0; liveU.insert(0);
	nU = 1; nL = 0;
	plan.clear();
	for (int i = 0; i < n; ++i) {
		StagePlan s;
		s.last = i == n - 1;
		s.uIn = uPhys[i];
		int out;
		if (s.last && n >= 2) out = 0;                // in place over U^0 (only read pointwise by its own cell's thread)
		else { out = 0; while (liveU.count(out)) ++out; }
		uPhys[i + 1] = out; liveU.insert(out); s.uOut = out;
		if (out + 1 > nU) nU = out + 1;
		bool neededLater = false;
		for (int m = i + 1; m < n; ++m) neededLater = neededLater || b(m, i) != 0;
		s.betaSelf = b(i, i);
		s.computeL = s.betaSelf != 0 || neededLater;
		s.lOut = -1;
		if (neededLater) {
			int l = 0; while (liveL.count(l)) ++l;
			lPhys[i] = l; liveL.insert(l); s.lOut = l;
			if (l + 1 > nL) nL = l + 1;
		}
		for (int k = 0; k <= i; ++k) if (a(i, k) != 0) s.alpha.push_back(Term{uPhys[k], a(i, k)});
		for (int k = 0; k < i; ++k) if (b(i, k) != 0) s.beta.push_back(Term{lPhys[k], b(i, k)});
		s.readsU = 1; s.readsL = (int)s.beta.size();
		for (auto& t : s.alpha) if (t.k != s.uIn) s.readsU++;
		plan.push_back(s);
		// release what no later stage reads (U^0 is kept: the last stage writes the new state there)
		for (int k = 1; k <= i; ++k) {
			bool later = false;
			for (int m = i + 1; m < n; ++m) later = later || a(m, k) != 0;
			if (!later && uPhys[k] > 0 && uPhys[k] != uPhys[i + 1]) liveU.erase(uPhys[k]);
		}
		for (int k = 0; k <= i; ++k) {
			bool later = false;
			for (int m = i + 1; m < n; ++m) later = later || b(m, k) != 0;
			if (!later && lPhys[k] >= 0) liveL.erase(lPhys[k]);
		}
	}
}

// Classic RK4-type tableaux (every stage but the last combines U^0 and its own L only; the last one sums alpha U^0 + dt sum_k beta_k L^k):
// the reference forms that sum term by term in ascending k (rk.lua:96-112), so its partial sums can be carried from stage to stage --
// stage k adds (beta_k dt) L^k while L^k is still in registers -- and the last stage reads ONE operand instead of U^0 and every earlier
// L: no L buffer is written at all, every stage has at most two staged operands (all of them fit the tallest marching tile), and the
// operations and their order are the reference's, so the result is bit-identical.  Needs a kernel that writes StageP::Aout.
static bool foldFinalStage(std::vector<StagePlan>& plan, int& nU, int& nL) {
	int const n = (int)plan.size();
	if (n < 3) return false;
	StagePlan const& last = plan[n - 1];
	if (last.alpha.size() != 1 || last.alpha[0].k != 0 || (int)last.beta.size() != n - 1 || !last.computeL) return false;
	for (int i = 0; i < n - 1; ++i) {
		StagePlan const& s = plan[i];
		if (!s.beta.empty() || s.lOut < 0 || !s.computeL || s.uOut == 0) return false;
		for (auto& t : s.alpha) if (t.k != 0) return false;
		if (last.beta[i].k != s.lOut) return false;                 // beta terms in ascending k: term i is L^i
	}
	if (plan[0].uIn != 0) return false;
	int const A = nU++;                                             // the running sum's buffer
	for (int i = 0; i < n - 1; ++i) {
		StagePlan& s = plan[i];
		s.accOut = A;
		s.accIn = i == 0 ? -1 : A;
		s.accCoef = i == 0 ? last.alpha[0].coef : 1.;
		s.accBetaSelf = last.beta[i].coef;
		s.lOut = -1;
		if (i > 0) s.readsU++;
	}
	StagePlan& l = plan[n - 1];
	l.alpha.assign(1, Term{A, 1.});
	l.beta.clear();
	l.readsU = 2; l.readsL = 0;
	nL = 0;
	return true;
}

struct FvBase {
	virtual ~FvBase() {}
	virtual int init() = 0;
	virtual int setState(const double* aos) = 0;
	virtual int getState(double* aos) = 0;
	virtual int setStateAsync(const double* aos) = 0;
	virtual int getStateAsync(double* aos) = 0;
	virtual int waitTransfers() = 0;
	virtual int stateDevPtr(void** p, long long* sy, long long* sz, long long* sv) = 0;
	virtual int boundary() = 0;
	virtual int setFixedBoundary(int face, const double* U, int n) = 0;
	virtual int addOp(const hb_op_desc* op, int* index) = 0;
	virtual int opsReset() = 0;
	virtual int opInfo(int op, int* iters, double* residual) = 0;
	virtual int constrainU() = 0;
	virtual int calcDT(double* out) = 0;
	virtual int step(double dt) = 0;
	virtual int update(int nsteps) = 0;
	virtual int getTime(double* t, double* dt) = 0;
	virtual int setTime(double t) = 0;
	virtual int calcDeriv(double dt, double* aos) = 0;
	virtual int describe(char* out, size_t cap) = 0;
	virtual int profile(int enable) = 0;
	virtual int profileRead(double* ms, long long* n) = 0;
	virtual int initDerivs() = 0;
	virtual int commInit(int nranks, int rank, const char* id) = 0;
	virtual int commDestroy() = 0;
	int nS = 0, nI = 0, nW = 0;
	long long cells = 0;
	long long launches = 0;
};

template<class real> struct Fv : FvBase {
	hb_ctx* ctx;
	hb_fv_desc d;
	const FvOps<real>* ops;
	GridP<real> grid;
	BcP bc;
	// ops (hydro/op): self-gravity, NoDiv over the Jacobi relaxation (hb_ops_kernels.cuh)
	struct OpState { hb_op_desc d; int pot, vec; OpCtl* ctl; };
	std::vector<OpState> opsV;
	real* opWrite = nullptr;               // writeBuf of relaxation.lua:52-57 (one variable)
	double* opPartial = nullptr;           // per-block partial sums / maxima
	int opBlocks = 0;
	bool opCtaRows = true;
	bool hasGrav = false, hasNoDiv = false;
	// corner-transport-upwind variant (hb_ctu_kernels.cuh): the reference's ULR and flux buffers
	bool useCTU = false;
	real* ctuULR = nullptr; real* ctuFlux = nullptr;
	bool seqBc = false;                    // a linear / quadratic / fixed face: ghost fill = the reference's x, y, z passes (fill_ghosts_axis)
	double* fixedDev = nullptr;            // [6][HB_FIXED_STRIDE] states of the 'fixed' faces
	std::vector<real*> upool, lpool;       // element (i=0,j=0,k=0) of variable 0; the allocation starts padX elements earlier
	std::vector<CUtensorMap> umaps;        // TMA descriptor of every U buffer (marching kernel)
	std::map<int, std::vector<CUtensorMap>> umapSets;   // the same for the configurations single stages run with (StagePlan::marchCfg): other box shapes
	CUtensorMap* opMapsDev = nullptr;                   // [stage][2 * HB_MAX_TERMS] operand tensor maps of the stages whose configuration stages the RK operands by TMA
	int padX = 0;                          // leading pad of every row: interior cell i=2 sits on a 128-byte boundary
	long long vstride = 0;                 // elements between variables (pitchX * S1 * S2)
	bool useMarch = false;
	int marchCfg = 0, marchBox[4] = {0, 0, 0, 0}, marchInfoV[7] = {0, 0, 0, 0, 0, 0, 0}, marchMaxOps = 0;
	bool rkFolded = false;         // the plan carries the last stage's running sum (foldFinalStage)
	real* scratchL = nullptr;
	real* opsScratch = nullptr;            // FvOps::scratchElems (ADM flux arrays)
	double* stagingAos = nullptr;
	// asynchronous state transfer (hb_fv_set_state_async / get_state_async): one staging buffer and one copy stream per direction,
	// so that an upload, the solver's work and a download overlap (PCIe is full duplex; the two directions use different copy engines)
	double* stagingIn = nullptr;
	double* stagingOut = nullptr;
	cudaStream_t upStream = nullptr, downStream = nullptr;
	cudaEvent_t evUpDone = nullptr, evInFree = nullptr, evOutReady = nullptr, evDownDone = nullptr;
	bool inUsed = false, downUsed = false;
	double* ctl = nullptr;                 // device: t, dt, cfl, fixedDT(<0 adaptive)
	unsigned long long* dtMinBits = nullptr;
	std::vector<StagePlan> plan;
	int nU = 0, nL = 0;
	bool dtValid = false, rkZeroed = false;
	int nonIntSync = 0;
	cudaGraphExec_t graphExec = nullptr;
	long long graphLaunches = 0;
	// optional per-launch timing of the stage kernel (CUDA events on the launching stream; eager mode only)
	bool profiling = false;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> profEvents;
	size_t profUsed = 0;
	// slab decomposition
	Nccl::comm_t comm = nullptr;
	int nranks = 1, rank = 0;
	int axis;
	// overlapped exchange (marching kernel, >= 3 chunks along the decomposed axis): the chunks next to the slab faces run first,
	// their ghost planes travel on commStream while the chunks in between are computed on the solver's stream
	cudaStream_t commStream = nullptr;
	cudaEvent_t evRim = nullptr, evXchg = nullptr;
	bool overlap = false;

	JitProgram* jit = nullptr;             // a run-time compiled equation (hb_fv_create_from_source): its FvOps launch from this program
	const FvOps<real>* OPS() const { tlsJit = jit; return ops; }
	Fv(hb_ctx* c, const hb_fv_desc& desc, const FvOps<real>* o, JitProgram* j = nullptr) : ctx(c), d(desc), ops(o), jit(j) {
		nS = o->nS; nI = o->nI; nW = o->nW;
		axis = d.dim - 1;
		ctxRetain(ctx);
	}
	~Fv() override {
		useDevice(ctx);
		cudaStreamSynchronize(ctx->stream);
		if (graphExec) cudaGraphExecDestroy(graphExec);
		for (auto& e : profEvents) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
		if (opMapsDev) cudaFree(opMapsDev);
		JitProgram* const jitToFree = jit;
		struct FreeJit { JitProgram* p; ~FreeJit() { if (p) jitFree(p); } } freeJit{jitToFree};
		for (auto p : upool) cudaFree(p - padX);
		for (auto p : lpool) cudaFree(p - padX);
		if (scratchL) cudaFree(scratchL - padX);
		if (opsScratch) cudaFree(opsScratch - padX);
		if (stagingAos) cudaFree(stagingAos);
		if (upStream) { cudaStreamSynchronize(upStream); cudaStreamDestroy(upStream); }
		if (downStream) { cudaStreamSynchronize(downStream); cudaStreamDestroy(downStream); }
		if (stagingIn) cudaFree(stagingIn);
		if (stagingOut) cudaFree(stagingOut);
		for (cudaEvent_t e : {evUpDone, evInFree, evOutReady, evDownDone}) if (e) cudaEventDestroy(e);
		if (ctl) cudaFree(ctl);
		if (fixedDev) cudaFree(fixedDev);
		if (ctuULR) cudaFree(ctuULR - padX);
		if (ctuFlux) cudaFree(ctuFlux - padX);
		if (opWrite) cudaFree(opWrite - padX);
		if (opPartial) cudaFree(opPartial);
		for (auto& o : opsV) if (o.ctl) cudaFree(o.ctl);
		if (dtMinBits) cudaFree(dtMinBits);
		if (comm) Nccl::get().CommDestroy(comm);
		if (evRim) cudaEventDestroy(evRim);
		if (evXchg) cudaEventDestroy(evXchg);
		if (commStream) cudaStreamDestroy(commStream);
		ctxRelease(ctx);
	}
	cudaStream_t st() const { return ctx->stream; }
	// allocation sizes: the variables' padded arrays plus one row of slack for the leading pad
	size_t uBytes() const { return sizeof(real) * ((size_t)nS * (size_t)vstride + (size_t)grid.strideY); }
	size_t lBytes() const { return sizeof(real) * ((size_t)nI * (size_t)vstride + (size_t)grid.strideY); }
	int allocPadded(real** out, size_t bytes) {
		real* p = nullptr;
		HB_CUDA(cudaMalloc(&p, bytes));
		HB_CUDA(cudaMemsetAsync(p, 0, bytes, st()));
		*out = p + padX;
		return HB_OK;
	}
	int encodeMap(real* U, CUtensorMap* map, const int* boxOverride = nullptr) {
		EncodeTiled_t enc = encodeTiledFn();
		if (!enc) return setError(HB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
		cuuint64_t dims[4] = {(cuuint64_t)grid.strideY, (cuuint64_t)grid.S[1], (cuuint64_t)grid.S[2], (cuuint64_t)nI};
		cuuint64_t strides[3] = {(cuuint64_t)grid.strideY * sizeof(real), (cuuint64_t)grid.strideZ * sizeof(real), (cuuint64_t)vstride * sizeof(real)};
		const int* bx = boxOverride ? boxOverride : marchBox;
		cuuint32_t box[4] = {(cuuint32_t)bx[0], (cuuint32_t)bx[1], (cuuint32_t)bx[2], (cuuint32_t)bx[3]};
		cuuint32_t es[4] = {1, 1, 1, 1};
		CUresult r = enc(map, sizeof(real) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)(U - padX),
			dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
			CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if (r != CUDA_SUCCESS) return setError(HB_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
		return HB_OK;
	}

	int init() override {
		useDevice(ctx);
		grid.dim = d.dim;
		for (int k = 0; k < 3; ++k) {
			grid.N[k] = k < d.dim ? d.n[k] : 1;
			grid.S[k] = k < d.dim ? d.n[k] + 2 * HB_G : 1;
			int const gn = k < d.dim ? d.global_n[k] : 1;
			double const dx = (d.maxs[k] - d.mins[k]) / double(gn);     // gridsolver.lua:406-409, host double then cast
			grid.dx[k] = real(dx);
			grid.invdx[k] = real(1.) / grid.dx[k];
		}
		// HBM layout: structure of arrays, rows padded so that the first interior cell of every row is 128-byte aligned
		// and the row pitch is a multiple of 128 bytes (coalesced warp stores; TMA needs 16-byte multiples)
		int const per128 = int(128 / sizeof(real));
		padX = per128 - HB_G;
		grid.strideY = (padX + grid.S[0] + per128 - 1) / per128 * per128;
		grid.strideZ = grid.strideY * grid.S[1];
		cells = (long long)grid.S[0] * grid.S[1] * grid.S[2];
		vstride = grid.strideZ * grid.S[2];
		grid.strideV = vstride;
		// fvsolver.cl:34-41,97-102: volume = prod dx (used axes), area_s = prod_{i != s} dx_i, scaled by 1/volume, in `real`
		real volume = 1;
		for (int k = 0; k < d.dim; ++k) volume = volume * grid.dx[k];
		grid.volOn = volume > real(1e-7);
		for (int s = 0; s < 3; ++s) {
			real area = 1;
			for (int k = 0; k < d.dim; ++k) if (k != s) area = area * grid.dx[k];
			grid.fluxOn[s] = s < d.dim && !(area <= real(1e-7));
			real a = area;
			if (a <= real(1e-7)) a = 0;
			grid.aov[s] = grid.volOn ? a * (real(1.) / volume) : real(0);
		}
		for (int k = 0; k < 6; ++k) bc.bc[k] = d.bc[k];
		bc.fixedState = nullptr;
		seqBc = false;
		// (a user-selected 'none' face, gridsolver.lua:618-621: the reference's later per-axis passes still fill the corner / edge ghost cells
		// next to it from the other axes' methods, which the composed single-pass source map skips -- run the per-axis sequence then.  Faces
		// owned by a neighbouring slab become 'none' in commInit, on `bc`, and do not count here: the exchange covers them.)
		for (int k = 0; k < 2 * d.dim; ++k) if (d.bc[k] >= HB_BC_LINEAR || d.bc[k] == HB_BC_NONE) seqBc = true;
		if (seqBc) {
			HB_CUDA(cudaMalloc(&fixedDev, sizeof(double) * 6 * HB_FIXED_STRIDE));
			HB_CUDA(cudaMemset(fixedDev, 0, sizeof(double) * 6 * HB_FIXED_STRIDE));
			bc.fixedState = fixedDev;
		}
		buildPlan(d.rk_order, d.alphas, d.betas, plan, nU, nL);
		for (auto& s : plan) {
			if ((int)s.alpha.size() > HB_MAX_TERMS || (int)s.beta.size() > HB_MAX_TERMS)
				return setError(HB_ERR_INVALID, "hb_fv_create: tableau row has more than 4 alpha or beta terms");
		}
		// stage kernel selection: the plane-marching TMA kernel where it is built (dim >= 2, 'plm cons' with minmod / superbee),
		// else the tile kernel.  d.stage_kernel: 0 auto, 1 tile kernel, 2 marching kernel (error if unavailable).
		bool const plm = d.use_plm != 0, flim = !plm && d.flux_limiter > 0;
		int cfg0 = 0;
		if (const char* e = getenv("HB_MARCH_CFG")) cfg0 = atoi(e);
		auto chooseKernel = [&]() -> int {
			// RK operands staged in shared memory per column thread: the largest count over the stages
			int maxOps = 0;
			for (auto& s : plan) if (s.operands() > maxOps) maxOps = s.operands();
			bool ok = false;
			if (d.stage_kernel != 1 && d.flux == HB_FLUX_ROE && d.use_plm == 1) {   // the marching kernel is built for Roe + 'plm cons'
				for (int pass = 0; pass < 2 && !ok; ++pass)
					for (int cfg = pass == 0 ? cfg0 : 0; !ok && OPS()->marchInfo(d.dim, plm, flim, d.slope_limiter, cfg, marchBox, marchInfoV); ++cfg) {
						size_t const smem = (size_t)marchInfoV[4] + sizeof(real) * (size_t)nI * (size_t)maxOps * (size_t)marchInfoV[5];
						if (marchInfoV[6] & 1) continue;   // configurations with the self-gravity epilogue are taken by hb_fv_add_op only
						if (smem <= 232448 - 1024) { ok = true; marchCfg = cfg; }
					}
			}
			// the general marching configurations (3-D): the other slope limiters, no reconstruction, Roe with a flux limiter, HLL / Rusanov / HLLC
			// with or without 'plm cons' -- everything else of these rows that round 1 ran through the tile kernel ($HB_MARCH_GEN=0: keep it there)
			const char* mg = getenv("HB_MARCH_GEN");
			// MEASURED AND NOT THE DEFAULT (profiles/r02k_gen_vs_tile.txt, r02j_bench_C4FL*.json): these modes run the LITERAL device functions
			// (three eigensystems per interface for the flux limiter, IEEE divisions and square roots behind slow-path branches), and with
			// that arithmetic the marching structure at 9 warps x 168 registers is 1.3-2.3 x SLOWER than the tile kernel at its higher occupancy
			// (3-D PLM + HLL 22.0 vs 17.3 ms per 256^3 stage, flux limiter 18.0 vs 9.8, 2-D PLM + HLL 2.9 vs 1.3): what these paths lack is
			// production-form arithmetic, not data movement.  So they stay on the tile kernel; $HB_MARCH_GEN=1 (2: also the flux-limiter mode)
			// or stage_kernel = 2 selects the marching kernel (kept as a second implementation: bit-identical in the strict build).
			bool const flimMode = d.use_plm == 0 && d.flux_limiter > 0;
			int const mgv = mg ? atoi(mg) : 0;
			if (!ok && d.stage_kernel != 1 && d.use_plm <= 1 && !d.use_ctu && OPS()->marchInfoGen && (d.stage_kernel == 2 || (mgv >= 1 && (!flimMode || mgv >= 2)))) {
				for (int cfg = kMarchGenBase; !ok && OPS()->marchInfoGen(d.dim, cfg, marchBox, marchInfoV); ++cfg) {
					size_t const smem = (size_t)marchInfoV[4] + sizeof(real) * (size_t)nI * (size_t)maxOps * (size_t)marchInfoV[5];
					if (smem <= 232448 - 1024) { ok = true; marchCfg = cfg; }
				}
			}
			if (d.stage_kernel == 2 && !ok) return setError(HB_ERR_INVALID, "hb_fv_create: the marching kernel is not built for this configuration");
			if (jit && !ok) return setError(HB_ERR_INVALID, "hb_fv_create_from_source: a run-time equation runs the marching kernels only (dim 2 or 3, Roe flux, usePLM = 'plm cons', slopeLimiter minmod or superbee, RK operands within shared memory)");
			useMarch = ok;
			marchMaxOps = maxOps;
			return HB_OK;
		};
		// the last stage's combination carried from stage to stage (foldFinalStage) where the stage kernel writes the running sum: the
		// marching kernels with the term-list epilogue (fv_march3, fv_march2d)
		{
			std::vector<StagePlan> const planStd = plan;
			int const nUStd = nU, nLStd = nL;
			const char* fe = getenv("HB_RK_FOLD");
			// measured (profiles/r02o_sweep_fold.txt): Euler 512 x 512 x 128 RK4 1.990 -> 1.945 ms per stage (every stage on the 32 x 15 tile); ideal
			// MHD 1.015 -> 1.044 (eight variables: the extra operand of the middle stages costs more than the last stage gains) -- so the default
			// folds for equations of up to five integrated variables; $HB_RK_FOLD=1 / 0 forces it on / off
			bool folded = (fe ? atoi(fe) != 0 : nI <= 5) && OPS()->eqnId != HB_EQN_ADM3D && !d.use_ctu && d.dim >= 2 && foldFinalStage(plan, nU, nL);
			if (int r = chooseKernel()) return r;
			if (folded && !(useMarch && ((marchInfoV[6] & 2) || (d.dim == 2 && marchBox[0] <= 40)))) {
				plan = planStd; nU = nUStd; nL = nLStd; folded = false;
				if (int r = chooseKernel()) return r;
			}
			rkFolded = folded;
		}
		if (nU < 2 && d.rk_order < 2) nU = 2;
		for (int k = 0; k < nU; ++k) {
			real* p = nullptr;
			if (int r = allocPadded(&p, uBytes())) return r;
			upool.push_back(p);
		}
		for (int k = 0; k < nL; ++k) {
			real* p = nullptr;
			if (int r = allocPadded(&p, lBytes())) return r;
			lpool.push_back(p);
		}
		{
			if (useMarch) {
				umaps.resize(nU);
				for (int k = 0; k < nU; ++k) if (int r = encodeMap(upool[k], &umaps[k])) return r;
				// per-stage configuration: the shared-memory budget is set by the stage's own operand count, so a stage with fewer RK operands
				// than the largest (classic RK4: 0, 1, 1, 4) takes the first (= preferred) configuration that fits IT.  Same planes per CTA as the
				// solver's configuration (the overlapped slab exchange selects chunks by that number); $HB_MARCH_PER_STAGE=0 switches it off.
				const char* ps = getenv("HB_MARCH_PER_STAGE");
				if ((!ps || atoi(ps) != 0) && marchCfg < kMarchGenBase) {      // (the general configurations are not in marchInfo's list)
					int const saveBox[4] = {marchBox[0], marchBox[1], marchBox[2], marchBox[3]};
					for (auto& s : plan) {
						int const n = s.operands();
						int box[4], info[7];
						for (int cfg = cfg0; cfg < marchCfg && OPS()->marchInfo(d.dim, plm, flim, d.slope_limiter, cfg, box, info); ++cfg) {
							size_t const smem = (size_t)info[4] + sizeof(real) * (size_t)nI * (size_t)n * (size_t)info[5];
							if ((info[6] & 1) || info[2] != marchInfoV[2] || smem > 232448 - 1024) continue;
							s.marchCfg = cfg;
							if (!umapSets.count(cfg)) {
								memcpy(marchBox, box, sizeof(box));
								std::vector<CUtensorMap>& v = umapSets[cfg];
								v.resize(nU);
								for (int k = 0; k < nU; ++k) if (int r = encodeMap(upool[k], &v[k])) return r;
								memcpy(marchBox, saveBox, sizeof(saveBox));
							}
							break;
						}
					}
				}
			}
		}
		if (int r = buildOpMaps()) return r;
		if (d.use_ctu) {
			useCTU = true;
			useMarch = false;
			size_t const block = (size_t)nI * (size_t)vstride;
			if (int r = allocPadded(&ctuULR, sizeof(real) * (2 * (size_t)d.dim * block + (size_t)grid.strideY))) return r;
			if (int r = allocPadded(&ctuFlux, sizeof(real) * ((size_t)d.dim * block + (size_t)grid.strideY))) return r;
		}
		if (OPS()->scratchElems) {
			if (int r = allocPadded(&opsScratch, sizeof(real) * ((size_t)OPS()->scratchElems(grid) + (size_t)grid.strideY))) return r;
		}
		HB_CUDA(cudaMalloc(&ctl, 4 * sizeof(double)));
		HB_CUDA(cudaMalloc(&dtMinBits, sizeof(unsigned long long)));
		double h[4] = {0., 0., d.cfl, d.use_fixed_dt ? d.fixed_dt : -1.};
		HB_CUDA(cudaMemcpyAsync(ctl, h, sizeof(h), cudaMemcpyHostToDevice, st()));
		reset_dtmin<<<1, 1, 0, st()>>>(dtMinBits);
		HB_CUDA(cudaGetLastError());
		HB_CUDA(cudaStreamSynchronize(st()));
		return HB_OK;
	}

	int ensureStaging() {
		if (!stagingAos) HB_CUDA(cudaMalloc(&stagingAos, sizeof(double) * (size_t)nS * (size_t)cells));
		return HB_OK;
	}
	void invalidateGraph() {
		if (graphExec) { cudaStreamSynchronize(st()); cudaGraphExecDestroy(graphExec); graphExec = nullptr; }
	}

	int setState(const double* aos) override {
		if (!aos) return setError(HB_ERR_INVALID, "hb_fv_set_state: null pointer");
		useDevice(ctx);
		if (int r = ensureStaging()) return r;
		size_t const n = (size_t)nS * (size_t)cells;
		HB_CUDA(cudaMemcpyAsync(stagingAos, aos, sizeof(double) * n, cudaMemcpyHostToDevice, st()));
		aos_to_soa<real><<<(unsigned)((n + 255) / 256), 256, 0, st()>>>(grid, nS, stagingAos, upool[0]);
		HB_CUDA(cudaGetLastError());
		launches++;
		dtValid = false; rkZeroed = false; nonIntSync = 2;
		return HB_OK;
	}
	int getState(double* aos) override {
		if (!aos) return setError(HB_ERR_INVALID, "hb_fv_get_state: null pointer");
		useDevice(ctx);
		if (int r = ensureStaging()) return r;
		size_t const n = (size_t)nS * (size_t)cells;
		soa_to_aos<real><<<(unsigned)((n + 255) / 256), 256, 0, st()>>>(grid, nS, upool[0], stagingAos);
		HB_CUDA(cudaGetLastError());
		launches++;
		HB_CUDA(cudaMemcpyAsync(aos, stagingAos, sizeof(double) * n, cudaMemcpyDeviceToHost, st()));
		HB_CUDA(cudaStreamSynchronize(st()));
		return HB_OK;
	}
	int ensureAsync() {
		if (upStream) return HB_OK;
		size_t const bytes = sizeof(double) * (size_t)nS * (size_t)cells;
		HB_CUDA(cudaMalloc(&stagingIn, bytes));
		HB_CUDA(cudaMalloc(&stagingOut, bytes));
		HB_CUDA(cudaStreamCreateWithFlags(&upStream, cudaStreamNonBlocking));
		HB_CUDA(cudaStreamCreateWithFlags(&downStream, cudaStreamNonBlocking));
		for (cudaEvent_t* e : {&evUpDone, &evInFree, &evOutReady, &evDownDone}) HB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
		return HB_OK;
	}
	// UBufObj:fromCPU without blocking the host: the copy runs on the upload stream, the solver's stream waits for it before the
	// AoS -> SoA kernel.  The host buffer (pinned) must stay unchanged until hb_fv_wait_transfers or a later blocking call.
	int setStateAsync(const double* aos) override {
		if (!aos) return setError(HB_ERR_INVALID, "hb_fv_set_state_async: null pointer");
		useDevice(ctx);
		if (int r = ensureAsync()) return r;
		size_t const n = (size_t)nS * (size_t)cells;
		if (inUsed) HB_CUDA(cudaStreamWaitEvent(upStream, evInFree, 0));     // the previous AoS -> SoA kernel has read the staging buffer
		HB_CUDA(cudaMemcpyAsync(stagingIn, aos, sizeof(double) * n, cudaMemcpyHostToDevice, upStream));
		HB_CUDA(cudaEventRecord(evUpDone, upStream));
		HB_CUDA(cudaStreamWaitEvent(st(), evUpDone, 0));
		aos_to_soa<real><<<(unsigned)((n + 255) / 256), 256, 0, st()>>>(grid, nS, stagingIn, upool[0]);
		HB_CUDA(cudaGetLastError());
		HB_CUDA(cudaEventRecord(evInFree, st()));
		inUsed = true;
		launches++;
		dtValid = false; rkZeroed = false; nonIntSync = 2;
		return HB_OK;
	}
	// UBufObj:toCPU without blocking the host: SoA -> AoS on the solver's stream, the copy on the download stream.
	int getStateAsync(double* aos) override {
		if (!aos) return setError(HB_ERR_INVALID, "hb_fv_get_state_async: null pointer");
		useDevice(ctx);
		if (int r = ensureAsync()) return r;
		size_t const n = (size_t)nS * (size_t)cells;
		if (downUsed) HB_CUDA(cudaStreamWaitEvent(st(), evDownDone, 0));     // the previous download has left the staging buffer
		soa_to_aos<real><<<(unsigned)((n + 255) / 256), 256, 0, st()>>>(grid, nS, upool[0], stagingOut);
		HB_CUDA(cudaGetLastError());
		launches++;
		HB_CUDA(cudaEventRecord(evOutReady, st()));
		HB_CUDA(cudaStreamWaitEvent(downStream, evOutReady, 0));
		HB_CUDA(cudaMemcpyAsync(aos, stagingOut, sizeof(double) * n, cudaMemcpyDeviceToHost, downStream));
		HB_CUDA(cudaEventRecord(evDownDone, downStream));
		downUsed = true;
		return HB_OK;
	}
	int waitTransfers() override {
		useDevice(ctx);
		if (upStream) HB_CUDA(cudaStreamSynchronize(upStream));
		HB_CUDA(cudaStreamSynchronize(st()));
		if (downStream) HB_CUDA(cudaStreamSynchronize(downStream));
		return HB_OK;
	}
	int stateDevPtr(void** p, long long* sy, long long* sz, long long* sv) override {
		if (p) *p = upool[0];
		if (sy) *sy = grid.strideY;
		if (sz) *sz = grid.strideZ;
		if (sv) *sv = grid.strideV;
		return HB_OK;
	}

	// ghost planes of the decomposed axis: 2 planes x (everything faster) are contiguous per variable
	int exchange(real* U, int nVars, cudaStream_t xs = nullptr) {
		if (!comm) return HB_OK;
		if (!xs) xs = st();
		Nccl& N = Nccl::get();
		long long const strideA = axis == 0 ? 1 : (axis == 1 ? grid.strideY : grid.strideZ);
		size_t const chunk = (size_t)(HB_G * strideA);
		int const S = grid.S[axis];
		bool const periodic = d.bc[2 * axis] == HB_BC_PERIODIC && d.bc[2 * axis + 1] == HB_BC_PERIODIC;
		int lo = rank - 1, hi = rank + 1;
		if (lo < 0) lo = periodic ? nranks - 1 : -1;
		if (hi >= nranks) hi = periodic ? 0 : -1;
		int const dtype = sizeof(real) == 8 ? kNcclFloat64 : kNcclFloat32;
		// Message order per peer: sends [low planes -> lo, high planes -> hi], receives [high ghosts <- hi, low ghosts <- lo].
		// With two ranks and a periodic axis lo == hi: NCCL pairs the sends and receives of one peer in issue order, and this
		// order makes my low planes land in the peer's high ghosts and my high planes in its low ghosts.
		HB_NCCL(N.GroupStart());
		for (int q = 0; q < nVars; ++q) {
			real* base = U + (size_t)q * grid.strideV;
			if (lo >= 0) HB_NCCL(N.Send(base + (size_t)HB_G * strideA, chunk, dtype, lo, comm, xs));
			if (hi >= 0) HB_NCCL(N.Send(base + (size_t)(S - 2 * HB_G) * strideA, chunk, dtype, hi, comm, xs));
		}
		for (int q = 0; q < nVars; ++q) {
			real* base = U + (size_t)q * grid.strideV;
			if (hi >= 0) HB_NCCL(N.Recv(base + (size_t)(S - HB_G) * strideA, chunk, dtype, hi, comm, xs));
			if (lo >= 0) HB_NCCL(N.Recv(base, chunk, dtype, lo, comm, xs));
		}
		HB_NCCL(N.GroupEnd());
		return HB_OK;
	}
	int reduceDtMin() {
		if (!comm) return HB_OK;
		// positive doubles order like their bit patterns, so the scalar can be min-reduced as a double
		HB_NCCL(Nccl::get().AllReduce(dtMinBits, dtMinBits, 1, kNcclFloat64, kNcclMin, comm, st()));
		return HB_OK;
	}

	int fillGhosts(real* U, int nVars, const BcP* methods = nullptr) {
		BcP const& b = methods ? *methods : bc;
		if (seqBc) {
			for (int a = 0; a < d.dim; ++a) { HB_CUDA(OPS()->ghosts(grid, b, U, nVars, -2 - a, false, st())); launches++; }
			return exchange(U, nVars);
		}
		HB_CUDA(OPS()->ghosts(grid, b, U, nVars, -1, false, st()));
		launches++;
		return exchange(U, nVars);
	}

	// ---- ops: hydro/op/relaxation.lua, selfgrav.lua, nodiv.lua
	int addOp(const hb_op_desc* o, int* index) override {
		if (!o) return setError(HB_ERR_INVALID, "hb_fv_add_op: null argument");
		if (!OPS()->opKernel) return setError(HB_ERR_INVALID, "hb_fv_add_op: ops are built for euler and mhd");
		if (useCTU && o->kind == HB_OP_SELFGRAV) return setError(HB_ERR_INVALID, "hb_fv_add_op: self-gravity is not built for the CTU variant");
		if (o->max_iters < 0) return setError(HB_ERR_INVALID, "hb_fv_add_op: max_iters < 0");
		OpState s; s.d = *o; s.ctl = nullptr; s.vec = -1;
		if (o->kind == HB_OP_SELFGRAV) s.pot = nS - 1;                                   // ePot: last variable of euler and mhd
		else if (o->kind == HB_OP_NODIV && d.eqn == HB_EQN_MHD) { s.pot = 8; s.vec = 5; }   // mhd: B = 5..7, psi = 8
		else return setError(HB_ERR_INVALID, "hb_fv_add_op: unknown op for this equation (selfgrav: euler, mhd; NoDiv: mhd)");
		useDevice(ctx);
		if (!opWrite) {
			if (int r = allocPadded(&opWrite, sizeof(real) * ((size_t)vstride + (size_t)grid.strideY))) return r;
			long long const rows = (long long)grid.S[1] * grid.S[2];
			long long const rowBlocks = (rows + HB_OP_WARPS - 1) / HB_OP_WARPS;
			// many short rows (3-D): a row per warp; few long rows (2-D, 1-D): a row per CTA.  At most 8 CTAs per SM, rows dealt round-robin.
			opCtaRows = !(grid.S[0] < 1024 && rowBlocks >= 148 * 4);
			long long const nb = opCtaRows ? rows : rowBlocks;
			opBlocks = (int)(nb < 148 * 8 ? nb : 148 * 8);
			HB_CUDA(cudaMalloc(&opPartial, sizeof(double) * (size_t)opBlocks));
		}
		HB_CUDA(cudaMalloc(&s.ctl, sizeof(OpCtl)));
		HB_CUDA(cudaMemset(s.ctl, 0, sizeof(OpCtl)));
		opsV.push_back(s);
		if (o->kind == HB_OP_SELFGRAV) {
			// the gravity source joins L in the stage kernel's epilogue: the marching configuration built with it (same tile geometry, so the
			// tensor maps stay valid), else the tile kernel
			hasGrav = true;
			if (useMarch && marchCfg >= kMarchGenBase) {                      // the general configurations have no gravity epilogue: tile kernel
				if (rkFolded) return setError(HB_ERR_INVALID, "hb_fv_add_op: the general marching configuration has no gravity epilogue; create the solver with $HB_RK_FOLD=0");
				useMarch = false;
			}
			if (useMarch) {
				bool const plm = d.use_plm != 0, flim = !plm && d.flux_limiter > 0;
				int box[4], info[7];
				bool found = false;
				for (int cfg = 0; !found && OPS()->marchInfo(d.dim, plm, flim, d.slope_limiter, cfg, box, info); ++cfg) {
					size_t const smem = (size_t)info[4] + sizeof(real) * (size_t)nI * (size_t)marchMaxOps * (size_t)info[5];
					if ((info[6] & 1) && (info[6] & 2) == (marchInfoV[6] & 2) && !memcmp(box, marchBox, sizeof(box)) && smem <= 232448 - 1024) { found = true; marchCfg = cfg; memcpy(marchInfoV, info, sizeof(info)); }
				}
				useMarch = found;
				if (!useMarch && rkFolded) return setError(HB_ERR_INVALID, "hb_fv_add_op: no marching configuration with the gravity epilogue for this solver; create it with $HB_RK_FOLD=0");
				for (auto& sPlan : plan) { sPlan.marchCfg = -1; sPlan.opTma = false; }   // one configuration (the one with the gravity epilogue) for every stage
			}
		}
		else hasNoDiv = true;
		invalidateGraph();
		dtValid = false;
		if (index) *index = (int)opsV.size() - 1;
		return HB_OK;
	}
	OpP<real> opParams(OpState const& s, real* U, int iter = 0) {
		OpP<real> p;
		memset(&p, 0, sizeof(p));
		p.kind = s.d.kind; p.U = U; p.writeBuf = opWrite; p.partial = opPartial; p.ctl = s.ctl; p.pot = s.pot; p.vec = s.vec;
		p.param = s.d.param; p.stopEpsilon = s.d.stop_epsilon; p.stopOnEpsilon = s.d.stop_on_epsilon; p.iter = iter;
		double v = 1; for (int k = 0; k < d.dim; ++k) v *= (double)d.global_n[k];     // the whole grid's interior (solver.volumeWithoutBorder)
		p.deferDecision = comm ? 1 : 0;
		p.volumeWithoutBorder = v; p.nBlocks = opBlocks; p.ctaRows = opCtaRows ? 1 : 0;
		// sweep 1 reads the potential in U and writes writeBuf, sweep 2 the other way round, ...
		real* potU = U + (size_t)s.pot * vstride;
		p.potIn = (iter & 1) ? potU : opWrite;
		p.potOut = (iter & 1) ? opWrite : potU;
		// poisson_jacobi.cl:57-117 on a cartesian grid, operation for operation, in `real`
		real volume = 1;
		for (int k = 0; k < d.dim; ++k) volume = volume * grid.dx[k];
		real const volL = real(.5) * (volume + volume), volR = real(.5) * (volume + volume), volAtX = volume;
		real diag = 0;
		for (int k = 0; k < d.dim; ++k) {
			real const dx = grid.dx[k];
			p.cS[k] = volR / (dx * dx);
			diag = diag - (volR + volL) / (dx * dx);
			p.sDiv[k] = real(.5 / double(dx));                 // nodiv.lua:104
			p.sGrad[k] = real(1. / (2. * double(dx)));         // nodiv.lua:150
		}
		p.invVol = real(1.) / volAtX;
		p.diag = diag / volAtX;
		p.invDiag = real(1.) / p.diag;
		return p;
	}
	int opLaunch(int which, OpP<real> const& p) { HB_CUDA(OPS()->opKernel(which, grid, p, st())); launches++; return HB_OK; }
	// Relaxation:potentialBoundary (relaxation.lua:135-150,198-200): the solver's boundary methods on the potential alone
	// A 'fixed' face writes whole states (its fixedCode knows nothing of args.fields): the field-restricted pass leaves it alone.
	int potentialBoundary(real* pot) {
		BcP b = bc;
		for (int k = 0; k < 6; ++k) if (b.bc[k] == HB_BC_FIXED) b.bc[k] = HB_BC_NONE;
		return fillGhosts(pot, 1, &b);
	}
	// Relaxation:relax (relaxation.lua:165-196); the stop test stays on the device (OpCtl::done), the two copies of the potential alternate
	int relax(OpState const& s, real* U) {
		if (int r = opLaunch(HB_OPK_BEGIN, opParams(s, U))) return r;
		for (int it = 1; it <= s.d.max_iters; ++it) {
			OpP<real> const p = opParams(s, U, it);
			if (int r = opLaunch(HB_OPK_JACOBI, p)) return r;
			if (int r = potentialBoundary(p.potOut)) return r;    // + the slab exchange of the potential's ghost planes
			if (comm) {
				// the residual is a sum over the whole grid: all-reduce the slabs' sums, then finish the iteration on every rank alike
				double* sum = reinterpret_cast<double*>(reinterpret_cast<char*>(s.ctl) + offsetof(OpCtl, sumLocal));
				if (s.d.stop_on_epsilon) HB_NCCL(Nccl::get().AllReduce(sum, sum, 1, kNcclFloat64, kNcclSum, comm, st()));
				if (int r = opLaunch(HB_OPK_DECIDE, p)) return r;
			}
		}
		return opLaunch(HB_OPK_FINAL_COPY, opParams(s, U));
	}
	int offsetPotential(OpState const& s, real* U) {       // selfgrav.lua:123-147
		OpP<real> const p = opParams(s, U);
		if (int r = opLaunch(HB_OPK_MAX, p)) return r;
		if (comm) {
			double* mx = reinterpret_cast<double*>(reinterpret_cast<char*>(s.ctl) + offsetof(OpCtl, maxVal));
			HB_NCCL(Nccl::get().AllReduce(mx, mx, 1, kNcclFloat64, kNcclMax, comm, st()));
		}
		return opLaunch(HB_OPK_OFFSET, p);
	}
	int opsReset() override {                              // solverbase.lua:2106-2111, relaxation.lua:152-158, selfgrav.lua:93-101
		useDevice(ctx);
		for (auto& s : opsV) {
			if (int r = opLaunch(HB_OPK_INIT, opParams(s, upool[0]))) return r;
			if (int r = potentialBoundary(upool[0] + (size_t)s.pot * vstride)) return r;
			if (int r = relax(s, upool[0])) return r;
			if (s.d.kind == HB_OP_SELFGRAV) if (int r = offsetPotential(s, upool[0])) return r;
			if (int r = fillGhosts(upool[0], nS)) return r;
		}
		dtValid = false;
		return HB_OK;
	}
	// op:step of every op after the integrator (solverbase.lua:3230-3237): boundary(), constrainU(), NoDiv:step (nodiv.lua:180-184)
	int opsStep() {
		for (auto& s : opsV) {
			if (s.d.kind != HB_OP_NODIV) continue;
			if (int r = fillGhosts(upool[0], nS)) return r;
			HB_CUDA(OPS()->constrainAll(grid, d.eqn_params, upool[0], st()));
			launches++;
			if (int r = fillGhosts(upool[0], nS)) return r;
			if (int r = relax(s, upool[0])) return r;
			if (int r = opLaunch(HB_OPK_NODIV, opParams(s, upool[0]))) return r;
		}
		return HB_OK;
	}
	int opInfo(int op, int* iters, double* residual) override {
		if (op < 0 || op >= (int)opsV.size()) return setError(HB_ERR_INVALID, "hb_fv_op_info: no such op");
		useDevice(ctx);
		OpCtl h;
		HB_CUDA(cudaMemcpyAsync(&h, opsV[op].ctl, sizeof(h), cudaMemcpyDeviceToHost, st()));
		HB_CUDA(cudaStreamSynchronize(st()));
		if (iters) *iters = h.lastIter;
		if (residual) *residual = h.lastResidual;
		return HB_OK;
	}

	int setFixedBoundary(int face, const double* U, int n) override {
		if (face < 0 || face >= 2 * d.dim || !U || n < 1 || n > nS) return setError(HB_ERR_INVALID, "hb_fv_set_fixed_boundary: bad face or state size");
		if (d.bc[face] != HB_BC_FIXED) return setError(HB_ERR_INVALID, "hb_fv_set_fixed_boundary: that face's boundary method is not 'fixed'");
		useDevice(ctx);
		HB_CUDA(cudaMemcpyAsync(fixedDev + (size_t)face * HB_FIXED_STRIDE, U, sizeof(double) * n, cudaMemcpyHostToDevice, st()));
		HB_CUDA(cudaStreamSynchronize(st()));
		return HB_OK;
	}
	int boundary() override {
		useDevice(ctx);
		return fillGhosts(upool[0], nS);
	}
	int constrainU() override {
		useDevice(ctx);
		HB_CUDA(OPS()->constrainAll(grid, d.eqn_params, upool[0], st()));
		launches++;
		dtValid = false;
		return fillGhosts(upool[0], nS);
	}
	int launchCalcDT() {
		reset_dtmin<<<1, 1, 0, st()>>>(dtMinBits);
		HB_CUDA(cudaGetLastError());
		HB_CUDA(OPS()->calcDT(grid, d.eqn_params, upool[0], dtMinBits, st()));
		launches += 2;
		if (int r = reduceDtMin()) return r;
		dtValid = true;
		return HB_OK;
	}
	int calcDT(double* out) override {
		if (!out) return setError(HB_ERR_INVALID, "hb_fv_calc_dt: null pointer");
		if (d.use_fixed_dt) { *out = d.fixed_dt; return HB_OK; }   // solverbase.lua:3007-3008
		useDevice(ctx);
		if (!dtValid) if (int r = launchCalcDT()) return r;
		unsigned long long bits;
		HB_CUDA(cudaMemcpyAsync(&bits, dtMinBits, 8, cudaMemcpyDeviceToHost, st()));
		HB_CUDA(cudaStreamSynchronize(st()));
		double m; memcpy(&m, &bits, 8);
		*out = d.cfl * m;                                           // solverbase.lua:3016
		return HB_OK;
	}

	void fillStageP(StageP<real>& sp, StagePlan const& s, bool wantDtMin) {
		memset(&sp, 0, sizeof(sp));
		sp.Uin = upool[s.uIn];
		sp.Uout = upool[s.uOut];
		sp.Lout = s.lOut >= 0 ? lpool[s.lOut] : nullptr;
		sp.nA = (int)s.alpha.size();
		for (int k = 0; k < sp.nA; ++k) {
			sp.aPtr[k] = upool[s.alpha[k].k]; sp.aCoef[k] = s.alpha[k].coef;
			if (s.alpha[k].k == s.uIn) sp.aOwnMask |= 1 << k;
		}
		sp.nB = (int)s.beta.size();
		for (int k = 0; k < sp.nB; ++k) { sp.bPtr[k] = lpool[s.beta[k].k]; sp.bCoef[k] = s.beta[k].coef; }
		for (int k = 0; k < sp.nA; ++k) {
			bool const own = (sp.aOwnMask >> k) & 1;
			sp.tCoef[sp.nT] = sp.aCoef[k];
			sp.tSlot[sp.nT++] = own ? -1 : sp.nOps;
			if (!own) sp.opPtr[sp.nOps++] = sp.aPtr[k];
		}
		for (int k = 0; k < sp.nB; ++k) {
			sp.tCoef[sp.nT] = sp.bCoef[k];
			sp.tBetaMask |= 1 << sp.nT;
			sp.tSlot[sp.nT++] = sp.nOps;
			sp.opPtr[sp.nOps++] = sp.bPtr[k];
		}
		if (s.accOut >= 0) {
			sp.Aout = upool[s.accOut];
			sp.accCoef = s.accCoef; sp.accBetaSelf = s.accBetaSelf;
			sp.accSlot = -1;
			if (s.accIn >= 0) { sp.accSlot = sp.nOps; sp.opPtr[sp.nOps++] = upool[s.accIn]; }
		}
		sp.betaSelf = s.betaSelf;
		sp.computeL = s.computeL ? 1 : 0;
		sp.dt = ctl + 1;
		sp.dtMinBits = wantDtMin ? dtMinBits : nullptr;
		sp.slopeLimiter = d.slope_limiter;
		sp.fluxLimiter = d.flux_limiter;
		sp.scratch = opsScratch;
		sp.flux = d.flux;
		sp.fluxParam = d.flux_param;
		sp.plmMode = d.use_plm;
		sp.opMaps = s.opTma && opMapsDev ? (const void*)(opMapsDev + (&s - &plan[0]) * 2 * HB_MAX_TERMS) : nullptr;
	}


	// FiniteVolumeSolver:calcDeriv with useCTU (fvsolver.lua:225-302) + the integrator's combination: calcLR, calcFlux, updateCTU,
	// boundaryLR, calcFlux, calcDerivFromFlux -- the reference's kernel sequence on its ULR / flux buffers (hb_ctu_kernels.cuh)
	int ctuStage(StageP<real> const& sp) {
		CtuP<real> c;
		memset(&c, 0, sizeof(c));
		c.ULR = ctuULR; c.flux = ctuFlux; c.blockStride = (long long)nI * vstride;
		real volume = 1;
		for (int k = 0; k < d.dim; ++k) volume = volume * grid.dx[k];
		c.invVolume = real(1.) / volume;
		for (int k = 0; k < d.dim; ++k) {
			real const volume_int = real(.5) * (volume + volume);
			c.areaL[k] = volume_int / grid.dx[k];
			c.areaR[k] = volume_int / grid.dx[k];
		}
		int n = 0;
		if (sp.computeL) {
			HB_CUDA(OPS()->ctuKernel(HB_CTUK_LR, grid, sp, c, d.eqn_params, st())); ++n;
			HB_CUDA(OPS()->ctuKernel(HB_CTUK_FLUX, grid, sp, c, d.eqn_params, st())); ++n;
			HB_CUDA(OPS()->ctuKernel(HB_CTUK_UPDATE, grid, sp, c, d.eqn_params, st())); ++n;
			// boundaryLR (gridsolver.lua:463-473,1241-1268): the solver's boundary methods on every L / R record; the reflected variables are the
			// same per record, so each block is filled like a state of nI variables
			long long const before = launches;
			for (int b = 0; b < 2 * d.dim; ++b) if (int r = fillGhosts(ctuULR + (size_t)b * (size_t)c.blockStride, nI)) return r;
			n += (int)(launches - before); launches = before;
			HB_CUDA(OPS()->ctuKernel(HB_CTUK_FLUX, grid, sp, c, d.eqn_params, st())); ++n;
		}
		HB_CUDA(OPS()->ctuKernel(HB_CTUK_FINISH, grid, sp, c, d.eqn_params, st())); ++n;
		tlsStageLaunches = n;
		return HB_OK;
	}

	// tensor maps over the RK operand buffers of every stage whose marching configuration fetches them by TMA (box = the tile without halo)
	int buildOpMaps() {
		for (auto& s : plan) s.opTma = false;
		if (!useMarch || marchCfg >= kMarchGenBase) return HB_OK;
		bool const plm = d.use_plm != 0, flim = !plm && d.flux_limiter > 0;
		std::vector<CUtensorMap> host(plan.size() * 2 * HB_MAX_TERMS);
		memset(host.data(), 0, host.size() * sizeof(CUtensorMap));
		bool any = false;
		for (size_t i = 0; i < plan.size(); ++i) {
			auto& s = plan[i];
			int box[4], info[7];
			if (!OPS()->marchInfo(d.dim, plm, flim, d.slope_limiter, stageCfg(s), box, info) || !(info[6] & 8)) continue;
			int const opBox[4] = {info[0], info[1], 1, nI};
			int n = 0;
			for (auto& t : s.alpha) if (t.k != s.uIn) { if (int r = encodeMap(upool[t.k], &host[i * 2 * HB_MAX_TERMS + n], opBox)) return r; ++n; }
			for (auto& t : s.beta) { if (int r = encodeMap(lpool[t.k], &host[i * 2 * HB_MAX_TERMS + n], opBox)) return r; ++n; }
			if (s.accOut >= 0 && s.accIn >= 0) { if (int r = encodeMap(upool[s.accIn], &host[i * 2 * HB_MAX_TERMS + n], opBox)) return r; ++n; }   // (fillStageP's operand order)
			s.opTma = true;
			any = true;
		}
		if (!any) return HB_OK;
		if (!opMapsDev) HB_CUDA(cudaMalloc(&opMapsDev, host.size() * sizeof(CUtensorMap)));
		HB_CUDA(cudaMemcpyAsync(opMapsDev, host.data(), host.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, st()));
		HB_CUDA(cudaStreamSynchronize(st()));
		return HB_OK;
	}
	int stageCfg(StagePlan const& s) const { return s.marchCfg >= 0 ? s.marchCfg : marchCfg; }
	const CUtensorMap* stageMap(StagePlan const& s) { return s.marchCfg >= 0 ? &umapSets[s.marchCfg][s.uIn] : &umaps[s.uIn]; }

	// integrator:integrate(dt, calcDeriv) + boundary/constrainU after every stage (rk.lua:91-165, fe.lua:33-49)
	int runStages() {
		bool const rk = d.rk_order >= 1;
		if (rk && !rkZeroed && nS > nI) {
			// rk.lua:94 clears UBuf (all fields, ghosts too); multAdd restores only the integrated fields (App. C #3)
			// (with ops the potentials of the current state come from op:resetState and are kept for the first relax)
			if (opsV.empty()) HB_CUDA(cudaMemsetAsync(upool[0] + (size_t)nI * vstride, 0, sizeof(real) * (size_t)(nS - nI) * vstride, st()));
			for (int k = 1; k < nU; ++k)
				HB_CUDA(cudaMemsetAsync(upool[k] + (size_t)nI * vstride, 0, sizeof(real) * (size_t)(nS - nI) * vstride, st()));
			rkZeroed = true;
		}
		if (!rk && nonIntSync > 0 && nS > nI) {
			// forward Euler keeps the non-integrated fields; carry them into the ping-pong partner
			HB_CUDA(cudaMemcpyAsync(upool[1] + (size_t)nI * vstride, upool[0] + (size_t)nI * vstride,
				sizeof(real) * (size_t)(nS - nI) * vstride, cudaMemcpyDeviceToDevice, st()));
			nonIntSync--;
		}
		bool const plm = d.use_plm != 0;
		bool const flim = !plm && d.flux_limiter > 0;
		for (auto& s : plan) {
			StageP<real> sp;
			fillStageP(sp, s, s.last && !hasNoDiv);
			if (hasGrav && s.computeL) {
				// op:addSource (solverbase.lua:3219-3223) = SelfGrav:addSource (selfgrav.lua:112-121): relax on the stage's input state, the
				// gravity source joins L inside the stage kernel, offsetPotential afterwards
				for (auto& o : opsV) if (o.d.kind == HB_OP_SELFGRAV) {
					if (int r = relax(o, upool[s.uIn])) return r;
					sp.gravPot = upool[s.uIn] + (size_t)o.pot * vstride;
				}
			}
			cudaEvent_t e0 = nullptr, e1 = nullptr;
			if (profiling) {
				if (profUsed == profEvents.size()) {
					cudaEvent_t a, b;
					HB_CUDA(cudaEventCreate(&a)); HB_CUDA(cudaEventCreate(&b));
					profEvents.push_back(std::make_pair(a, b));
				}
				e0 = profEvents[profUsed].first; e1 = profEvents[profUsed].second; profUsed++;
				HB_CUDA(cudaEventRecord(e0, st()));
			}
			if (useMarch && overlap && opsV.empty()) {   // (with ops the exchange stays in-stream)
				int const nv = rk ? nI : nS;
				HB_CUDA(OPS()->march(d.dim, d.slope_limiter, stageCfg(s), stageMap(s), padX, grid, sp, d.eqn_params, 1, st()));
				HB_CUDA(OPS()->ghosts(grid, bc, upool[s.uOut], nv, axis, true, st()));
				HB_CUDA(cudaEventRecord(evRim, st()));
				HB_CUDA(cudaStreamWaitEvent(commStream, evRim, 0));
				if (int r = exchange(upool[s.uOut], nv, commStream)) return r;
				HB_CUDA(cudaEventRecord(evXchg, commStream));
				HB_CUDA(OPS()->march(d.dim, d.slope_limiter, stageCfg(s), stageMap(s), padX, grid, sp, d.eqn_params, 2, st()));
				if (profiling) HB_CUDA(cudaEventRecord(e1, st()));
				HB_CUDA(OPS()->ghosts(grid, bc, upool[s.uOut], nv, axis, false, st()));
				HB_CUDA(cudaStreamWaitEvent(st(), evXchg, 0));
				launches += 4;
				continue;
			}
			tlsStageLaunches = 1;
			if (useCTU) { if (int r = ctuStage(sp)) return r; }
			else if (useMarch) HB_CUDA(OPS()->march(d.dim, d.slope_limiter, stageCfg(s), stageMap(s), padX, grid, sp, d.eqn_params, 0, st()));
			else HB_CUDA(OPS()->stage(d.dim, plm, flim, grid, sp, d.eqn_params, st()));
			if (profiling) HB_CUDA(cudaEventRecord(e1, st()));
			launches += tlsStageLaunches;
			if (!opsV.empty()) {
				if (hasGrav && s.computeL) for (auto& o : opsV) if (o.d.kind == HB_OP_SELFGRAV) if (int r = offsetPotential(o, upool[s.uIn])) return r;
				if (nS > nI && s.uOut != s.uIn) {
					// the fields the integrator does not touch: rk.lua:94 leaves them 0 in the stage's result, fe.lua keeps them
					real* dst = upool[s.uOut] + (size_t)nI * vstride;
					if (rk) HB_CUDA(cudaMemsetAsync(dst, 0, sizeof(real) * (size_t)(nS - nI) * vstride, st()));
					else HB_CUDA(cudaMemcpyAsync(dst, upool[s.uIn] + (size_t)nI * vstride, sizeof(real) * (size_t)(nS - nI) * vstride, cudaMemcpyDeviceToDevice, st()));
				}
			}
			if (int r = fillGhosts(upool[s.uOut], rk ? nI : nS)) return r;
		}
		if (plan.back().uOut != 0) {   // order <= 1: ping-pong
			std::swap(upool[0], upool[plan.back().uOut]);
			if (useMarch) {
				std::swap(umaps[0], umaps[plan.back().uOut]);
				for (auto& kv : umapSets) std::swap(kv.second[0], kv.second[plan.back().uOut]);
			}
		}
		if (hasNoDiv) {
			// NoDiv changes B after the last stage: the CFL reduction fused into that stage would be stale
			if (int r = opsStep()) return r;
			if (int r = fillGhosts(upool[0], nS)) return r;   // SolverBase:update ends with boundary() (solverbase.lua:3186); not redundant after an op:step
			reset_dtmin<<<1, 1, 0, st()>>>(dtMinBits);
			HB_CUDA(cudaGetLastError());
			HB_CUDA(OPS()->calcDT(grid, d.eqn_params, upool[0], dtMinBits, st()));
			launches += 2;
		}
		if (int r = reduceDtMin()) return r;
		dtValid = true;
		return HB_OK;
	}

	int step(double dt) override {
		useDevice(ctx);
		set_step<<<1, 1, 0, st()>>>(ctl, dtMinBits, dt);
		HB_CUDA(cudaGetLastError());
		launches++;
		return runStages();
	}

	int oneUpdate() {
		begin_step<<<1, 1, 0, st()>>>(ctl, dtMinBits);
		HB_CUDA(cudaGetLastError());
		launches++;
		return runStages();
	}

	int update(int nsteps) override {
		useDevice(ctx);
		if (nsteps <= 0) return HB_OK;
		if (!d.use_fixed_dt && !dtValid) if (int r = launchCalcDT()) return r;
		bool const graphable = d.use_graph && d.rk_order >= 2 && !profiling;
		int done = 0;
		if (graphable) {
			if (d.rk_order >= 1 && !rkZeroed) { if (int r = oneUpdate()) return r; done = 1; }   // one-off memsets stay outside the graph
			if (done < nsteps && !graphExec) {
				long long const before = launches;
				cudaGraph_t graph = nullptr;
				HB_CUDA(cudaStreamBeginCapture(st(), cudaStreamCaptureModeThreadLocal));
				int r = oneUpdate();
				cudaError_t e = cudaStreamEndCapture(st(), &graph);
				if (r) { if (graph) cudaGraphDestroy(graph); return r; }
				if (e != cudaSuccess) return cudaFail(e, "cudaStreamEndCapture");
				graphLaunches = launches - before;
				launches = before;
				e = cudaGraphInstantiate(&graphExec, graph, 0);
				cudaGraphDestroy(graph);
				if (e != cudaSuccess) return cudaFail(e, "cudaGraphInstantiate");
			}
			for (; done < nsteps; ++done) {
				HB_CUDA(cudaGraphLaunch(graphExec, st()));
				launches += graphLaunches;
			}
			return HB_OK;
		}
		for (; done < nsteps; ++done) if (int r = oneUpdate()) return r;
		return HB_OK;
	}

	int getTime(double* t, double* dt) override {
		useDevice(ctx);
		double h[2];
		HB_CUDA(cudaMemcpyAsync(h, ctl, sizeof(h), cudaMemcpyDeviceToHost, st()));
		HB_CUDA(cudaStreamSynchronize(st()));
		if (t) *t = h[0];
		if (dt) *dt = h[1];
		return HB_OK;
	}
	int setTime(double t) override {
		useDevice(ctx);
		HB_CUDA(cudaMemcpyAsync(ctl, &t, sizeof(double), cudaMemcpyHostToDevice, st()));
		HB_CUDA(cudaStreamSynchronize(st()));
		return HB_OK;
	}

	int calcDeriv(double dt, double* aos) override {
		if (!aos) return setError(HB_ERR_INVALID, "hb_fv_calc_deriv: null pointer");
		useDevice(ctx);
		if (int r = ensureStaging()) return r;
		if (!scratchL) { if (int r = allocPadded(&scratchL, uBytes())) return r; }
		else HB_CUDA(cudaMemsetAsync(scratchL - padX, 0, uBytes(), st()));
		// keep the device dt of a running simulation intact: use a private dt slot (ctl+1 is restored below)
		double saved[2];
		HB_CUDA(cudaMemcpyAsync(saved, ctl, sizeof(saved), cudaMemcpyDeviceToHost, st()));
		HB_CUDA(cudaStreamSynchronize(st()));
		HB_CUDA(cudaMemcpyAsync(ctl + 1, &dt, sizeof(double), cudaMemcpyHostToDevice, st()));
		StageP<real> sp;
		memset(&sp, 0, sizeof(sp));
		sp.Uin = upool[0]; sp.Uout = nullptr; sp.Lout = scratchL;
		sp.computeL = 1; sp.dt = ctl + 1;
		sp.slopeLimiter = d.slope_limiter; sp.fluxLimiter = d.flux_limiter; sp.scratch = opsScratch; sp.flux = d.flux; sp.fluxParam = d.flux_param; sp.plmMode = d.use_plm;
		bool const plm = d.use_plm != 0;
		if (useCTU) { if (int r = ctuStage(sp)) return r; }
		else if (useMarch) HB_CUDA(OPS()->march(d.dim, d.slope_limiter, marchCfg, &umaps[0], padX, grid, sp, d.eqn_params, 0, st()));
		else HB_CUDA(OPS()->stage(d.dim, plm, !plm && d.flux_limiter > 0, grid, sp, d.eqn_params, st()));
		launches++;
		HB_CUDA(cudaMemcpyAsync(ctl + 1, &saved[1], sizeof(double), cudaMemcpyHostToDevice, st()));
		size_t const n = (size_t)nS * (size_t)cells;
		soa_to_aos<real><<<(unsigned)((n + 255) / 256), 256, 0, st()>>>(grid, nS, scratchL, stagingAos);
		HB_CUDA(cudaGetLastError());
		launches++;
		HB_CUDA(cudaMemcpyAsync(aos, stagingAos, sizeof(double) * n, cudaMemcpyDeviceToHost, st()));
		HB_CUDA(cudaStreamSynchronize(st()));
		return HB_OK;
	}

	int describe(char* out, size_t cap) override {
		std::ostringstream o;
		int ti[5];
		bool const plm = d.use_plm != 0;
 		OPS()->tileInfo(d.dim, plm, !plm && d.flux_limiter > 0, ti);
		if (useMarch) for (int k = 0; k < 5; ++k) ti[k] = marchInfoV[k];
		o << "kernel=" << (useCTU ? "ctu(unfused:calcLR,calcFlux,updateCTU,boundaryLR,calcFlux,finish)" : useMarch ? (d.dim == 2 && marchBox[0] <= 40 ? "fv_march2d(warp-per-pencil,tma)" : ((marchInfoV[6] & 2) ? "fv_march3(tma,split-barrier)" : "fv_march(tma)")) : (OPS()->eqnId == HB_EQN_ADM3D ? "adm_flux_xyz+adm_update" : "fv_stage(tile)")) << " cfg=" << marchCfg << " pitchX=" << grid.strideY << " padX=" << padX << " ";
		o << "eqn=" << OPS()->eqnId << " real=" << sizeof(real) * 8 << " dim=" << d.dim << " strict_fp=" << d.strict_fp
		  << " tile=" << ti[0] << "x" << ti[1] << "x" << ti[2] << " threads=" << ti[3] << " smem=" << ti[4]
		  << " Ubufs=" << nU << " Lbufs=" << nL << (overlap ? " exchange=overlapped" : (comm ? " exchange=in-stream" : ""));
		if (rkFolded) o << " rkFold=1";
		if (useMarch) { o << " stageCfgs="; for (size_t i = 0; i < plan.size(); ++i) o << (i ? "," : "") << stageCfg(plan[i]); }
		o << "\n";
		int words = 0;
		for (size_t i = 0; i < plan.size(); ++i) {
			auto& s = plan[i];
			int const w = s.readsU + s.readsL + 1 + (s.lOut >= 0 ? 1 : 0) + (s.accOut >= 0 ? 1 : 0);
			words += w;
			o << "stage " << i << ": in=U" << s.uIn << " out=U" << s.uOut << " storeL=" << s.lOut << " alpha=" << s.alpha.size()
			  << " beta=" << s.beta.size() << " betaSelf=" << s.betaSelf << " words=" << w << "\n";
		}
		o << "words_per_cell_update=" << words << " (x nI=" << nI << " x " << sizeof(real) << " B)\n";
		snprintf(out, cap, "%s", o.str().c_str());
		return HB_OK;
	}

	int profile(int enable) override {
		useDevice(ctx);
		HB_CUDA(cudaStreamSynchronize(st()));
		profiling = enable != 0;
		profUsed = 0;
		return HB_OK;
	}
	int profileRead(double* ms, long long* n) override {
		useDevice(ctx);
		HB_CUDA(cudaStreamSynchronize(st()));
		double total = 0;
		for (size_t i = 0; i < profUsed; ++i) {
			float t = 0;
			HB_CUDA(cudaEventElapsedTime(&t, profEvents[i].first, profEvents[i].second));
			total += t;
		}
		if (ms) *ms = total;
		if (n) *n = (long long)profUsed;
		return HB_OK;
	}
	int initDerivs() override {
		useDevice(ctx);
		if (!OPS()->initDerivs) return HB_OK;      // equations without an initDerivs kernel (hydro/init/init.lua:231-235)
		HB_CUDA(OPS()->initDerivs(grid, upool[0], st()));
		launches++;
		dtValid = false;
		return HB_OK;
	}
	int commInit(int nr, int rk_, const char* id) override {
		useDevice(ctx);
		Nccl& N = Nccl::get();
		if (!N.ok) return setError(HB_ERR_CUDA, "hb_fv_comm_init: " + N.why);
		if (nr < 1 || rk_ < 0 || rk_ >= nr || !id) return setError(HB_ERR_INVALID, "hb_fv_comm_init: bad arguments");
		if (comm) return setError(HB_ERR_INVALID, "hb_fv_comm_init: already initialised");
		nranks = nr; rank = rk_;
		if (nr == 1) return HB_OK;
		Nccl::uid u; memcpy(u.internal, id, 128);
		HB_NCCL(N.CommInitRank(&comm, nr, u, rk_));
		// faces owned by a neighbouring slab are filled by exchange(), not by the local boundary kernel
		bool const periodic = d.bc[2 * axis] == HB_BC_PERIODIC && d.bc[2 * axis + 1] == HB_BC_PERIODIC;
		if (rank > 0 || periodic) bc.bc[2 * axis] = HB_BC_NONE;
		if (rank < nr - 1 || periodic) bc.bc[2 * axis + 1] = HB_BC_NONE;
		// overlap needs chunks that are neither first nor last along the decomposed (= marching) axis; HB_OVERLAP=0 switches it off
		const char* ov = getenv("HB_OVERLAP");
		int const km = marchInfoV[2] > 0 ? marchInfoV[2] : 1;
		// (kernels with the rim / interior split -- marchInfoV[6] bit 2 -- overlap whatever the chunk count: SURVEY 8e's 64 planes per GPU are one chunk)
		bool const rim = (marchInfoV[6] & 4) != 0;
		overlap = useMarch && !seqBc && (!ov || atoi(ov) != 0) && (rim ? grid.N[axis] >= 2 * HB_G + 1 : (grid.N[axis] + km - 1) / km >= 3);
		if (overlap) {
			HB_CUDA(cudaStreamCreateWithFlags(&commStream, cudaStreamNonBlocking));
			HB_CUDA(cudaEventCreateWithFlags(&evRim, cudaEventDisableTiming));
			HB_CUDA(cudaEventCreateWithFlags(&evXchg, cudaEventDisableTiming));
		}
		invalidateGraph();
		dtValid = false;
		return HB_OK;
	}
	int commDestroy() override {
		if (comm) { useDevice(ctx); cudaStreamSynchronize(st()); Nccl::get().CommDestroy(comm); comm = nullptr; }
		return HB_OK;
	}
};

}   // namespace hb

struct hb_fv { hb::FvBase* impl; };

using namespace hb;

extern "C" {

size_t hb_sizeof_fv_desc(void) { return sizeof(hb_fv_desc); }
size_t hb_sizeof_op_desc(void) { return sizeof(hb_op_desc); }
int hb_fv_create(hb_ctx* ctx, const hb_fv_desc* d, hb_fv** out) {
	if (!ctx || !d || !out) return setError(HB_ERR_INVALID, "hb_fv_create: null argument");
	*out = nullptr;
	if (d->dim < 1 || d->dim > 3) return setError(HB_ERR_INVALID, "hb_fv_create: dim must be 1..3");
	for (int k = 0; k < d->dim; ++k) {
		if (d->n[k] < 1 || d->global_n[k] < d->n[k]) return setError(HB_ERR_INVALID, "hb_fv_create: bad grid size");
		for (int m = 0; m < 2; ++m) if (d->bc[2 * k + m] < 0 || d->bc[2 * k + m] > HB_BC_FIXED) return setError(HB_ERR_INVALID, "hb_fv_create: unknown boundary method");
	}
	if (d->rk_order < 0 || d->rk_order > 4) return setError(HB_ERR_INVALID, "hb_fv_create: rk_order must be 0..4");
	if (d->use_plm < 0 || d->use_plm > 10) return setError(HB_ERR_INVALID, "hb_fv_create: usePLM must be none, 'plm cons', 'plm athena', 'plm prim', 'plm cons with flux', 'plm eig', 'plm eig prim' or 'plm eig prim ref' ('ppm' is not built)");
	if (d->use_plm >= 2 && d->eqn == HB_EQN_ADM3D) return setError(HB_ERR_INVALID, "hb_fv_create: 'plm athena' / 'plm prim' are built for euler and mhd");
	if (d->use_plm >= 2 && d->eqn == HB_EQN_MHD && d->dim == 3) return setError(HB_ERR_INVALID, "hb_fv_create: 'plm athena' / 'plm prim' for mhd are built for 1-D and 2-D grids (the 3-D tile does not hold both face states of 8 variables in shared memory)");
	if (d->slope_limiter < 0 || d->slope_limiter > 19 || d->flux_limiter < 0 || d->flux_limiter > 19) return setError(HB_ERR_INVALID, "hb_fv_create: limiter index out of range");
	if (d->use_plm && d->flux_limiter != 0) return setError(HB_ERR_INVALID, "hb_fv_create: usePLM requires fluxLimiter 'donor cell' (gridsolver.lua:119)");
	if (d->flux < 0 || d->flux > HB_FLUX_EULER_HLLC) return setError(HB_ERR_INVALID, "hb_fv_create: unknown flux");
	if (d->flux == HB_FLUX_EULER_HLLC && d->eqn != HB_EQN_EULER) return setError(HB_ERR_INVALID, "hb_fv_create: euler-hllc only works with the euler equation (euler-hllc.cl:8-10)");
	if (d->flux == HB_FLUX_EULER_HLLC && (d->flux_param < 0 || d->flux_param > 2)) return setError(HB_ERR_INVALID, "hb_fv_create: hllcMethod must be 0, 1 or 2");
	if (d->flux != HB_FLUX_ROE && d->flux_limiter != 0) return setError(HB_ERR_INVALID, "hb_fv_create: only the Roe flux uses a flux limiter (hydro/flux/roe.lua:5-19)");
	if (d->flux != HB_FLUX_ROE && d->eqn == HB_EQN_ADM3D) return setError(HB_ERR_INVALID, "hb_fv_create: hll / rusanov are built for euler and mhd");
	if (d->use_ctu) {
		if (d->use_plm != 1 || d->dim < 2) return setError(HB_ERR_INVALID, "hb_fv_create: useCTU is built for 'plm cons' in 2-D and 3-D (the reference switches it off in 1-D, gridsolver.lua:112-115)");
		if (d->eqn == HB_EQN_ADM3D) return setError(HB_ERR_INVALID, "hb_fv_create: useCTU is built for euler and mhd");
		if (d->stage_kernel == 2) return setError(HB_ERR_INVALID, "hb_fv_create: useCTU runs the unfused kernel sequence, not the marching kernel");
	}
	if (d->eqn == HB_EQN_ADM3D) {
		if (d->use_plm) return setError(HB_ERR_INVALID, "hb_fv_create: adm3d runs the Roe flux with a flux limiter on cell-centred states; usePLM is not built for it");
		for (int k = 0; k < 2 * d->dim; ++k) if (d->bc[k] == HB_BC_MIRROR) return setError(HB_ERR_INVALID, "hb_fv_create: mirror boundaries are not built for adm3d");
	}
	FvBase* impl = nullptr;
	bool const strict = d->strict_fp != 0;
	if (ctx->real_bytes == 8) {
		const FvOps<double>* o = nullptr;
		if (d->eqn == HB_EQN_EULER) o = strict ? ops_euler_f64_strict() : ops_euler_f64_fast();
		else if (d->eqn == HB_EQN_MHD) o = strict ? ops_mhd_f64_strict() : ops_mhd_f64_fast();
		else if (d->eqn == HB_EQN_ADM3D) o = strict ? ops_adm3d_f64_strict() : ops_adm3d_f64_fast();
		if (o) impl = new Fv<double>(ctx, *d, o);
	} else {
		const FvOps<float>* o = nullptr;
		if (d->eqn == HB_EQN_EULER) o = strict ? ops_euler_f32_strict() : ops_euler_f32_fast();
		else if (d->eqn == HB_EQN_MHD) o = strict ? ops_mhd_f32_strict() : ops_mhd_f32_fast();
		else if (d->eqn == HB_EQN_ADM3D) o = strict ? ops_adm3d_f32_strict() : ops_adm3d_f32_fast();
		if (o) impl = new Fv<float>(ctx, *d, o);
	}
	if (!impl) return setError(HB_ERR_INVALID, "hb_fv_create: unknown equation id");
	if (int r = impl->init()) { delete impl; return r; }
	*out = new hb_fv{impl};
	return HB_OK;
}
// An equation supplied as source (hb_jit.cu): `header_src` defines the class template `eqn_type`<real, FAST> with the plug-in contract of
// csrc/hb_eqn_euler.cuh; it is registered as the include file `header_name` (naming an embedded header replaces it).  desc->eqn is ignored;
// desc->eqn_params are handed to the equation's makeParams.  NVRTC compiles the marching / ghost / CFL / constrain kernels over it.
int hb_fv_create_from_source(hb_ctx* ctx, const hb_fv_desc* d, const char* header_name, const char* header_src, const char* eqn_type,
	hb_fv** out, char* log, size_t log_cap)
{
	if (log && log_cap) log[0] = 0;
	if (!ctx || !d || !out || !header_name || !header_src || !eqn_type) return setError(HB_ERR_INVALID, "hb_fv_create_from_source: null argument");
	*out = nullptr;
	if (d->dim < 2 || d->dim > 3) return setError(HB_ERR_INVALID, "hb_fv_create_from_source: dim must be 2 or 3 (the marching kernels)");
	for (int k = 0; k < d->dim; ++k) {
		if (d->n[k] < 1 || d->global_n[k] < d->n[k]) return setError(HB_ERR_INVALID, "hb_fv_create_from_source: bad grid size");
		for (int m = 0; m < 2; ++m) if (d->bc[2 * k + m] < 0 || d->bc[2 * k + m] > HB_BC_FIXED) return setError(HB_ERR_INVALID, "hb_fv_create_from_source: unknown boundary method");
	}
	if (d->rk_order < 0 || d->rk_order > 4) return setError(HB_ERR_INVALID, "hb_fv_create_from_source: rk_order must be 0..4");
	if (d->use_plm != 1 || d->flux != HB_FLUX_ROE || d->flux_limiter != 0 || d->use_ctu || (d->slope_limiter != 8 && d->slope_limiter != 18))
		return setError(HB_ERR_INVALID, "hb_fv_create_from_source: a run-time equation runs Roe + usePLM = 'plm cons' with the minmod or superbee slope limiter");
	std::string err, lg;
	bool const strict = d->strict_fp != 0;
	JitProgram* P = ctx->real_bytes == 8
		? jitCompile<double>(ctx, header_name, header_src, eqn_type, d->dim, d->slope_limiter, strict, d->eqn_params, err, lg)
		: jitCompile<float>(ctx, header_name, header_src, eqn_type, d->dim, d->slope_limiter, strict, d->eqn_params, err, lg);
	if (log && log_cap) snprintf(log, log_cap, "%s", lg.c_str());
	if (!P) return setError(HB_ERR_COMPILE, "hb_fv_create_from_source: " + err);
	FvBase* impl = ctx->real_bytes == 8 ? (FvBase*)new Fv<double>(ctx, *d, jitOps<double>(P), P) : (FvBase*)new Fv<float>(ctx, *d, jitOps<float>(P), P);
	if (int r = impl->init()) { delete impl; return r; }
	*out = new hb_fv{impl};
	return HB_OK;
}
int hb_fv_destroy(hb_fv* fv) { if (fv) { delete fv->impl; delete fv; } return HB_OK; }
#define HB_FV(fv) if (!(fv)) return setError(HB_ERR_INVALID, "null hb_fv handle")
int hb_fv_num_states(hb_fv* fv, int* ns, int* ni, int* nw) { HB_FV(fv); if (ns) *ns = fv->impl->nS; if (ni) *ni = fv->impl->nI; if (nw) *nw = fv->impl->nW; return HB_OK; }
long long hb_fv_num_cells(hb_fv* fv) { return fv ? fv->impl->cells : 0; }
int hb_fv_set_state(hb_fv* fv, const double* aos) { HB_FV(fv); return fv->impl->setState(aos); }
int hb_fv_get_state(hb_fv* fv, double* aos) { HB_FV(fv); return fv->impl->getState(aos); }
int hb_fv_set_state_async(hb_fv* fv, const double* aos) { HB_FV(fv); return fv->impl->setStateAsync(aos); }
int hb_fv_get_state_async(hb_fv* fv, double* aos) { HB_FV(fv); return fv->impl->getStateAsync(aos); }
int hb_fv_wait_transfers(hb_fv* fv) { HB_FV(fv); return fv->impl->waitTransfers(); }
int hb_fv_state_devptr(hb_fv* fv, void** p, long long* sy, long long* sz, long long* sv) { HB_FV(fv); return fv->impl->stateDevPtr(p, sy, sz, sv); }
int hb_fv_boundary(hb_fv* fv) { HB_FV(fv); return fv->impl->boundary(); }
int hb_fv_add_op(hb_fv* fv, const hb_op_desc* op, int* index) { HB_FV(fv); return fv->impl->addOp(op, index); }
int hb_fv_ops_reset(hb_fv* fv) { HB_FV(fv); return fv->impl->opsReset(); }
int hb_fv_op_info(hb_fv* fv, int op, int* it, double* res) { HB_FV(fv); return fv->impl->opInfo(op, it, res); }
int hb_fv_set_fixed_boundary(hb_fv* fv, int face, const double* cons, int n) { HB_FV(fv); return fv->impl->setFixedBoundary(face, cons, n); }
int hb_fv_constrainU(hb_fv* fv) { HB_FV(fv); return fv->impl->constrainU(); }
int hb_fv_init_derivs(hb_fv* fv) { HB_FV(fv); return fv->impl->initDerivs(); }
int hb_fv_calc_dt(hb_fv* fv, double* dt) { HB_FV(fv); return fv->impl->calcDT(dt); }
int hb_fv_step(hb_fv* fv, double dt) { HB_FV(fv); return fv->impl->step(dt); }
int hb_fv_update(hb_fv* fv, int nsteps) { HB_FV(fv); return fv->impl->update(nsteps); }
int hb_fv_get_time(hb_fv* fv, double* t, double* dt) { HB_FV(fv); return fv->impl->getTime(t, dt); }
int hb_fv_set_time(hb_fv* fv, double t) { HB_FV(fv); return fv->impl->setTime(t); }
int hb_fv_calc_deriv(hb_fv* fv, double dt, double* aos) { HB_FV(fv); return fv->impl->calcDeriv(dt, aos); }
int hb_fv_launch_count(hb_fv* fv, long long* n) { HB_FV(fv); if (n) *n = fv->impl->launches; return HB_OK; }
int hb_fv_describe(hb_fv* fv, char* out, size_t cap) { HB_FV(fv); if (!out || !cap) return setError(HB_ERR_INVALID, "hb_fv_describe: bad buffer"); return fv->impl->describe(out, cap); }
int hb_fv_profile(hb_fv* fv, int enable) { HB_FV(fv); return fv->impl->profile(enable); }
int hb_fv_profile_read(hb_fv* fv, double* ms, long long* n) { HB_FV(fv); return fv->impl->profileRead(ms, n); }
// host-only (no device needed): the static plan of one update for a tableau in the reference's (alpha, beta) form -- which physical buffer
// every stage reads and writes, its terms in evaluation order, and (fold != 0) the running sum of the last stage -- as text, one line per
// stage; tests/test_rk_plan.py executes it on scalars against the direct evaluation of rk.lua:91-165
int hb_rk_plan(int order, const double* alphas, const double* betas, int fold, char* out, size_t cap) {
	if (!out || !cap || order < 0 || order > 4 || (order >= 1 && (!alphas || !betas))) return setError(HB_ERR_INVALID, "hb_rk_plan: bad argument (order 0..4, order x order tables as in hb_fv_desc)");
	std::vector<StagePlan> plan;
	int nU = 0, nL = 0;
	buildPlan(order, alphas, betas, plan, nU, nL);
	bool const folded = fold && foldFinalStage(plan, nU, nL);
	std::ostringstream o;
	o.precision(17);
	o << "nU=" << nU << " nL=" << nL << " folded=" << (folded ? 1 : 0) << "\n";
	for (size_t i = 0; i < plan.size(); ++i) {
		StagePlan const& s = plan[i];
		o << "stage " << i << " in=" << s.uIn << " out=" << s.uOut << " lout=" << s.lOut << " computeL=" << (s.computeL ? 1 : 0) << " betaSelf=" << s.betaSelf
		  << " operands=" << s.operands() << " alpha=";
		for (size_t k = 0; k < s.alpha.size(); ++k) o << (k ? "," : "") << s.alpha[k].k << ":" << s.alpha[k].coef;
		o << " beta=";
		for (size_t k = 0; k < s.beta.size(); ++k) o << (k ? "," : "") << s.beta[k].k << ":" << s.beta[k].coef;
		o << " accOut=" << s.accOut << " accIn=" << s.accIn << " accCoef=" << s.accCoef << " accBetaSelf=" << s.accBetaSelf << "\n";
	}
	snprintf(out, cap, "%s", o.str().c_str());
	return HB_OK;
}
int hb_ghost_source(int j, int S, int bcMin, int bcMax, int* flip, int* skip) {
	bool f, s;
	int const r = ghostSource(j, S, bcMin, bcMax, f, s);
	if (flip) *flip = f; if (skip) *skip = s;
	return r;
}
// unit-test hook (not used by the product path): evaluate one device function per item on the GPU
int hb_debug_eval(hb_ctx* ctx, int eqn, int strict, int kind, int side, int n, const double* params, const double* aux4,
	const double* in, size_t inCount, double* out, size_t outCount)
{
	if (!ctx || !params || !aux4 || !in || !out || n <= 0) return setError(HB_ERR_INVALID, "hb_debug_eval: bad argument");
	useDevice(ctx);
	double *dIn = nullptr, *dOut = nullptr, *dAux = nullptr;
	HB_CUDA(cudaMalloc(&dIn, inCount * 8)); HB_CUDA(cudaMalloc(&dOut, outCount * 8)); HB_CUDA(cudaMalloc(&dAux, 4 * 8));
	HB_CUDA(cudaMemcpy(dIn, in, inCount * 8, cudaMemcpyHostToDevice));
	HB_CUDA(cudaMemcpy(dAux, aux4, 4 * 8, cudaMemcpyHostToDevice));
	cudaError_t e = cudaErrorInvalidValue;
	if (ctx->real_bytes == 8) {
		const FvOps<double>* o = eqn == HB_EQN_EULER ? (strict ? ops_euler_f64_strict() : ops_euler_f64_fast()) : (strict ? ops_mhd_f64_strict() : ops_mhd_f64_fast());
		e = o->debugEval(kind, side, n, params, dAux, dIn, dOut, ctx->stream);
	} else {
		const FvOps<float>* o = eqn == HB_EQN_EULER ? (strict ? ops_euler_f32_strict() : ops_euler_f32_fast()) : (strict ? ops_mhd_f32_strict() : ops_mhd_f32_fast());
		e = o->debugEval(kind, side, n, params, dAux, dIn, dOut, ctx->stream);
	}
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	if (e == cudaSuccess) e = cudaMemcpy(out, dOut, outCount * 8, cudaMemcpyDeviceToHost);
	cudaFree(dIn); cudaFree(dOut); cudaFree(dAux);
	if (e != cudaSuccess) return cudaFail(e, "hb_debug_eval");
	return HB_OK;
}
int hb_comm_unique_id(char* out128) {
	if (!out128) return setError(HB_ERR_INVALID, "hb_comm_unique_id: null pointer");
	Nccl& N = Nccl::get();
	if (!N.ok) return setError(HB_ERR_CUDA, "hb_comm_unique_id: " + N.why);
	Nccl::uid u;
	HB_NCCL(N.GetUniqueId(&u));
	memcpy(out128, u.internal, 128);
	return HB_OK;
}
int hb_fv_comm_init(hb_fv* fv, int nranks, int rank, const char* id) { HB_FV(fv); return fv->impl->commInit(nranks, rank, id); }
int hb_fv_comm_destroy(hb_fv* fv) { HB_FV(fv); return fv->impl->commDestroy(); }

}   // extern "C"
