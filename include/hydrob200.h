/* hydrob200.h -- C ABI of libhydrob200.so: the drop-in boundary between hydro-cl-lua's LuaJIT host side and the
 * B200 (sm_100a CUDA) backend.
 *
 * The reference reaches its device only through lua-opencl's `cl.obj.*` objects.  Each group
 * below names the reference call sites it replaces (file:line into the reference tree).  A LuaJIT binding
 * `ffi.cdef`s this header verbatim (INTEGRATION.md shows it); the Python host mirror in hydro-cl-lua_b200/
 * binds it through ctypes.  No C++ or torch types cross this boundary: plain pointers, sizes and PODs.
 *
 * Conventions: every function returns 0 on success and a non-zero code on failure; hb_last_error() returns the
 * message of the calling thread's last failure.  Objects are opaque handles.  A context owns one CUDA stream (the
 * analogue of the reference's single in-order command queue, hydro/solver/solverbase.lua:532-533); all work of a
 * context is stream-ordered and asynchronous unless stated.  Thread-compatible, not thread-safe: one host
 * thread per context, as in the reference (single LuaJIT thread).
 * There is no CPU fallback: without a CUDA device every device entry point fails with HB_ERR_NO_DEVICE.
 */
#ifndef HYDROB200_H
#define HYDROB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_OK 0
#define HB_ERR_INVALID 1      /* bad argument / unsupported configuration */
#define HB_ERR_NO_DEVICE 2    /* no CUDA device or driver */
#define HB_ERR_CUDA 3         /* CUDA runtime / driver / NVRTC / NCCL failure (see hb_last_error) */
#define HB_ERR_COMPILE 4      /* NVRTC compilation failed: log returned by hb_module_compile */

typedef struct hb_ctx hb_ctx;
typedef struct hb_buf hb_buf;
typedef struct hb_module hb_module;
typedef struct hb_kernel hb_kernel;
typedef struct hb_fv hb_fv;

const char* hb_last_error(void);
int hb_version(void);

/* ---- environment: replaces CLEnv{precision, cpu, ...} -> env.real, env.devices, env.cmds
 *      (hydro/app.lua:891-929; CL_DEVICE_MAX_WORK_GROUP_SIZE use at hydro/solver/solverbase.lua:765) */
int hb_device_count(int* count);
int hb_ctx_create(int device, int real_bytes /* 8 = double, 4 = float (hydro/app.lua:892) */, hb_ctx** out);
int hb_ctx_destroy(hb_ctx* ctx);
int hb_ctx_real_bytes(hb_ctx* ctx);
int hb_device_name(hb_ctx* ctx, char* out, size_t cap);
int hb_device_max_threads(hb_ctx* ctx, int* out);
int hb_device_sm_count(hb_ctx* ctx, int* out);
int hb_sync(hb_ctx* ctx);                                   /* cmds:finish(), hydro/solver/solverbase.lua:2096,2133 */
void* hb_ctx_stream(hb_ctx* ctx);                           /* the cudaStream_t, for event timing by the caller */
/* device-side timing on the context's stream (CUDA events) */
int hb_timer_start(hb_ctx* ctx);
int hb_timer_stop(hb_ctx* ctx, float* ms_out);              /* synchronises the stop event */
/* pinned host memory for hb_buf_write / hb_buf_read / hb_fv_set_state / hb_fv_get_state sources */
int hb_host_alloc(size_t bytes, void** out);
int hb_host_free(void* p);

/* ---- buffers: replaces CLBuffer{env,name,type,count}:fromCPU/toCPU/fill, ctx:buffer{rw,size}, enqueueCopyBuffer,
 *      clEnqueueCopyBufferRect (hydro/solver/solverbase.lua:1077-1095,1132,1417-1428; hydro/int/rk.lua:25,33,63;
 *      hydro/int/int.lua:6-8; hydro/solver/choppedup.lua:199-231) */
int hb_buf_alloc(hb_ctx* ctx, size_t bytes, hb_buf** out);
int hb_buf_free(hb_buf* buf);
size_t hb_buf_size(hb_buf* buf);
void* hb_buf_devptr(hb_buf* buf);
int hb_buf_write(hb_buf* buf, const void* host, size_t offset, size_t bytes);        /* fromCPU; async w.r.t. pinned host memory */
int hb_buf_read(hb_buf* buf, void* host, size_t offset, size_t bytes);               /* toCPU; blocking */
int hb_buf_fill(hb_buf* buf, const void* pattern, size_t pattern_bytes, size_t offset, size_t bytes);
int hb_buf_copy(hb_buf* dst, size_t dst_offset, hb_buf* src, size_t src_offset, size_t bytes);
/* origins/region in (bytes along x, rows, slices); pitches in bytes -- same meaning as clEnqueueCopyBufferRect */
int hb_buf_copy_rect(hb_buf* dst, hb_buf* src, const size_t src_origin[3], const size_t dst_origin[3], const size_t region[3],
	size_t src_row_pitch, size_t src_slice_pitch, size_t dst_row_pitch, size_t dst_slice_pitch);

/* ---- programs and kernels: replaces Program{name,code}:compile{buildOptions} + binary cache, program:kernel(name,...),
 *      k.obj:setArg(i,x), k(...) / cmds:enqueueNDRangeKernel{kernel,globalSize,localSize}
 *      (hydro/solver/solverbase.lua:558-713,1328-1345,1696-1699; hydro/solver/fvsolver.lua:216-221;
 *       hydro/solver/gridsolver.lua:1206-1209,1277-1309).  Source is CUDA C++ compiled by NVRTC for sm_100a. */
int hb_module_compile(hb_ctx* ctx, const char* cuda_src, const char* name, const char* const* opts, int nopts,
	hb_module** out, char* log, size_t log_cap);
int hb_module_free(hb_module* m);
int hb_kernel_get(hb_module* m, const char* name, hb_kernel** out);
int hb_kernel_set_arg(hb_kernel* k, int index, const void* value, size_t bytes);     /* by-value argument */
int hb_kernel_set_arg_buf(hb_kernel* k, int index, hb_buf* buf);                     /* device pointer argument */
int hb_kernel_launch(hb_kernel* k, const size_t global_size[3], const size_t local_size[3], size_t shared_bytes);
int hb_kernel_free(hb_kernel* k);                                                    /* before hb_module_free of its module */
/* The same for source in the OpenCL-C dialect the reference's templates emit (hydro/code/math.cl, hydro/eqn/cl/*.cl, hydro/solver/*.cl,
 * hydro/eqn/*.cl after template expansion; assembled at hydro/solver/solverbase.lua:1686-1699): `kernel` / `global` / `constant`
 * qualifiers, C99 compound literals `(real3){.x = a, ...}` (hydro/code/math.cl:47-52), get_global_id() etc. are rewritten / provided by
 * hb_cl_prelude(), unannotated functions compile as device functions.  hb_cl_translate is the rewrite alone (no device needed). */
const char* hb_cl_prelude(void);
int hb_cl_translate(const char* cl_src, char* out, size_t cap, size_t* needed);
int hb_module_compile_opencl(hb_ctx* ctx, const char* cl_src, const char* name, const char* const* opts, int nopts,
	hb_module** out, char* log, size_t log_cap);

/* ---- reductions: replaces env:reduce{count,op,buffer,...}() -> host scalar
 *      (hydro/solver/solverbase.lua:1350-1376, used at :3016) */
#define HB_REDUCE_MIN 0
#define HB_REDUCE_MAX 1
#define HB_REDUCE_SUM 2
int hb_reduce(hb_ctx* ctx, hb_buf* buf, size_t count, int op, double* host_out);     /* elements are the ctx's `real` */

/* ---- the fused finite-volume path (no single reference analogue: it is what FiniteVolumeSolver:calcDeriv,
 *      integrator:integrate, SolverBase:calcDT/step/update/boundary/constrainU enqueue together;
 *      hydro/solver/fvsolver.lua:225-302, hydro/int/fe.lua:33-49, hydro/int/rk.lua:47-167,
 *      hydro/solver/solverbase.lua:2116-2127,3004-3023,3026-3238, hydro/solver/gridsolver.lua:1272-1320) */
#define HB_EQN_EULER 0        /* hydro/eqn/euler.lua: eqn_params = { heatCapacityRatio, rhoMin, PMin } */
#define HB_EQN_MHD 1          /* hydro/eqn/mhd.lua:   eqn_params = { heatCapacityRatio, mu0 / unit_kg_m_per_C2 } */
#define HB_EQN_ADM3D 2        /* hydro/eqn/adm3d.lua (noZeroRowsInFlux, useShift 'none'): eqn_params = { f_eqn option index
                               * (hydro/eqn/einstein.lua:42-48, 0-based), a_convCoeff, d_convCoeff, V_convCoeff } */
#define HB_FLUX_ROE 0         /* hydro/flux/roe.cl:17-163 (usesFluxLimiter) */
#define HB_FLUX_HLL 1         /* hydro/flux/hll.cl:5-74, hllCalcWaveMethod 'Davis direct bounded' (hll.lua:10); Euler and MHD */
#define HB_FLUX_RUSANOV 2     /* hydro/flux/rusanov.cl:4-33; Euler and MHD */
#define HB_FLUX_EULER_HLLC 3  /* hydro/flux/euler-hllc.cl:14-243; Euler; flux_param = hllcMethod 0 | 1 | 2 (euler-hllc.lua:17) */
#define HB_BC_PERIODIC 0      /* hydro/solver/gridsolver.lua:638-651 */
#define HB_BC_MIRROR 1        /* :654-744 */
#define HB_BC_FREEFLOW 2      /* :766-780 */
#define HB_BC_NONE 3          /* :618-621; also: face owned by a neighbouring slab (hb_fv_exchange fills it) */
#define HB_BC_LINEAR 4        /* :782-813 linear extrapolation, every state variable */
#define HB_BC_QUADRATIC 5     /* :815-846 quadratic extrapolation */
#define HB_BC_FIXED 6         /* :746-764 Dirichlet: the state set with hb_fv_set_fixed_boundary (a fixedCode that does not depend on the
                               * cell, e.g. hydro/init/euler.lua:1859-1879 'square cavity' lid, hydro/eqn/einstein.lua:62-80 flat space) */

typedef struct hb_fv_desc {
	int eqn;                  /* HB_EQN_* */
	int dim;                  /* 1..3 */
	int n[3];                 /* interior cells of THIS rank's slab per axis (1 on unused axes) */
	int global_n[3];          /* interior cells of the whole grid (== n without decomposition); defines grid_dx */
	int use_plm;              /* 0 = none, 1 = 'plm cons' (hydro/solver/plm.cl:27-91), 2 = 'plm athena' (plm.cl:782-879; euler, mhd in 1-D / 2-D), 3 = 'plm athena'
	                           * with the face states assigned L = left, R = right: the order that reproduces the errors the reference
	                           * recorded for this scheme (its tree has them the other way round at plm.cl:877-878); 4 = 'plm prim' (plm.cl:191-253); 5 = 'plm cons with flux' (plm.cl:95-187);
	                           * 6 = 'plm eig' (plm.cl:256-427); 7 = 'plm eig prim', 8 = 'plm eig prim ref' (plm.cl:536-778: result->L is the state extrapolated
	                           * towards +side, as the tree has it); 9, 10 = 7, 8 with L and R exchanged.  2-10: euler, mhd in 1-D / 2-D; tile kernel.
	                           * 'piecewise constant' (plm.cl:10-24: L = R = U) is use_plm = 0 with flux_limiter = 0 */
	int slope_limiter;        /* 0-based index into hydro/app.lua:614-635 */
	int flux_limiter;         /* 0-based; 0 = 'donor cell' = no flux limiter (hydro/solver/fvsolver.lua:61-63) */
	int bc[6];                /* xmin,xmax,ymin,ymax,zmin,zmax: HB_BC_* */
	int rk_order;             /* 0 = forward Euler (hydro/int/fe.lua), else Butcher order (hydro/int/rk.lua) */
	double alphas[16];        /* row-major [order][order], hydro/int/all.lua */
	double betas[16];
	double mins[3], maxs[3];  /* whole-grid domain (solver_t mins/maxs) */
	double cfl;               /* hydro/solver/solverbase.lua:806, config.lua:37 */
	double fixed_dt;          /* used when use_fixed_dt (hydro/solver/solverbase.lua:3007-3008) */
	int use_fixed_dt;
	double eqn_params[16];
	int strict_fp;            /* 1: kernels built with -fmad=false (no FMA contraction); 0: production kernels */
	int use_graph;            /* 1: replay each update() as a captured CUDA graph */
	int stage_kernel;         /* 0: auto; 1: tile kernel (fv_stage); 2: plane-marching TMA kernel (fv_march), error if not built for the config */
	int flux;                 /* HB_FLUX_*: the solver's flux plug-in (hydro/flux/*.lua; solver.flux = 'roe' | 'hll' | 'rusanov' | 'euler-hllc') */
	int flux_param;           /* euler-hllc: solver.flux.hllcMethod */
	int use_ctu;              /* solver.useCTU (hydro/solver/gridsolver.lua:102-115, fvsolver.lua:246-272, ctu.cl): corner-transport-upwind correction of the
	                           * PLM face states between two flux passes; needs use_plm = 1 and dim >= 2 (the reference switches it off in 1-D); euler, mhd */
} hb_fv_desc;

size_t hb_sizeof_fv_desc(void);                              /* sizeof(hb_fv_desc), for bindings that mirror the struct */
int hb_fv_create(hb_ctx* ctx, const hb_fv_desc* desc, hb_fv** out);
int hb_fv_destroy(hb_fv* fv);
/* The codegen seam (hydro/eqn/eqn.lua:506-513,577-635 eqn:initCodeModules; hydro/solver/solverbase.lua:1686-1699): the equation's device
 * functions arrive as SOURCE at run time.  `header_src` defines the class template `eqn_type`<real, FAST> with the plug-in contract of
 * csrc/hb_eqn_euler.cuh (nS, nI, nW, eqnId, Params / makeParams, Eig, eigenForInterface, leftTransform, rightTransform, waves,
 * fluxFromCons, constrainU, calcDTCell, mirrorFlips); NVRTC instantiates the fused marching kernels (fv_march3 / fv_march2d) and the ghost
 * / CFL / constrain kernels over it for sm_100a.  `header_name` is the include name the source is registered under; naming one of the
 * library's own headers (e.g. "hb_eqn_euler.cuh") replaces it.  desc->eqn is ignored, desc->eqn_params go to makeParams.  Needs
 * dim 2 or 3, flux roe, use_plm 1, slope limiter minmod (8) or superbee (18). */
int hb_fv_create_from_source(hb_ctx* ctx, const hb_fv_desc* desc, const char* header_name, const char* header_src, const char* eqn_type,
	hb_fv** out, char* log, size_t log_cap);
int hb_fv_num_states(hb_fv* fv, int* num_states, int* num_int_states, int* num_waves);
long long hb_fv_num_cells(hb_fv* fv);                        /* ghost-inclusive cell count of this rank's slab */
/* state exchange in the reference's layout: AoS cons_t records of doubles, INDEX order (hydro/app.lua:976-984),
 * ghost cells included; converted to/from the SoA `real` device layout on the device */
int hb_fv_set_state(hb_fv* fv, const double* aos_host);
int hb_fv_get_state(hb_fv* fv, double* aos_host);           /* blocking */
/* The same transfers without blocking the host (clEnqueueWrite/ReadBuffer with blocking = false): upload and download run on their
 * own copy streams with their own staging buffers, ordered against the solver's stream by events, so that the upload of the next
 * problem, the update of the current one and the download of the previous one overlap.  Host buffers must be pinned (hb_host_alloc)
 * and stay untouched until hb_fv_wait_transfers returns. */
int hb_fv_set_state_async(hb_fv* fv, const double* aos_host);
int hb_fv_get_state_async(hb_fv* fv, double* aos_host);
int hb_fv_wait_transfers(hb_fv* fv);
int hb_fv_state_devptr(hb_fv* fv, void** soa_dev, long long* stride_y, long long* stride_z, long long* stride_var);
int hb_fv_set_fixed_boundary(hb_fv* fv, int face, const double* cons, int n);   /* face 0..5 = xmin,xmax,..,zmax; n <= numStates doubles; zeros until set */
/* ---- ops either side of the step (hydro/op; solver.ops, solverbase.lua:2106-2111, 3219-3237): self-gravity and NoDiv over the Jacobi
 *      Poisson relaxation (hydro/op/relaxation.lua, poisson.cl, poisson_jacobi.cl).  Once added, hb_fv_step / hb_fv_update run
 *      op:addSource inside every stage and op:step after the integrator, as SolverBase:step does.  On a slab-decomposed grid the potential's
 *      ghost planes travel with every sweep's ghost fill and the residual sum / potential maximum are all-reduced over NCCL. */
#define HB_OP_SELFGRAV 1      /* hydro/op/selfgrav.lua + selfgrav.cl; euler, mhd; potential = ePot; param = gravitationalConstant / unit_m3_per_kg_s2 */
#define HB_OP_NODIV 2         /* hydro/op/nodiv.lua with the Jacobi parent (noDivPoissonSolver=jacobi); mhd; potential = psi, vector = B */
typedef struct hb_op_desc {
	int kind;                 /* HB_OP_* */
	int max_iters;            /* Relaxation.maxIters, relaxation.lua:26 (20) */
	int stop_on_epsilon;      /* relaxation.lua:24 (true) */
	double stop_epsilon;      /* relaxation.lua:25 (1e-10) */
	double param;
} hb_op_desc;
size_t hb_sizeof_op_desc(void);                              /* sizeof(hb_op_desc), for bindings that mirror the struct */
int hb_fv_add_op(hb_fv* fv, const hb_op_desc* op, int* index_out);
int hb_fv_ops_reset(hb_fv* fv);                              /* op:resetState() + boundary() of every op (solverbase.lua:2106-2111); call after set_state / boundary */
int hb_fv_op_info(hb_fv* fv, int op, int* last_iter, double* last_residual);   /* Relaxation.lastIter / lastResidual (blocking) */
int hb_fv_boundary(hb_fv* fv);                               /* solver:boundary(), gridsolver.lua:1316 */
int hb_fv_init_derivs(hb_fv* fv);                            /* initDerivsKernelObj(), hydro/init/init.lua:231-235 (adm3d.cl:196-243); no-op for other equations */
int hb_fv_constrainU(hb_fv* fv);                             /* solver:constrainU() = kernel + boundary(), solverbase.lua:2116-2127 */
int hb_fv_calc_dt(hb_fv* fv, double* dt_out);                /* solver:calcDT(), solverbase.lua:3004-3023 (blocking) */
int hb_fv_step(hb_fv* fv, double dt);                        /* solver:step(dt), solverbase.lua:3193-3238 */
int hb_fv_update(hb_fv* fv, int nsteps);                     /* nsteps x solver:update(); dt stays on the device */
int hb_fv_get_time(hb_fv* fv, double* t_out, double* last_dt_out);   /* blocking read of t and the last dt */
int hb_fv_set_time(hb_fv* fv, double t);
int hb_fv_calc_deriv(hb_fv* fv, double dt, double* aos_host_out);    /* FiniteVolumeSolver:calcDeriv into a zeroed deriv buffer (blocking) */
int hb_fv_launch_count(hb_fv* fv, long long* kernel_launches);       /* kernels this object has launched so far */
int hb_fv_describe(hb_fv* fv, char* out, size_t cap);          /* text: kernel, tile shape, smem, per-stage plan (reads / writes per cell) */
int hb_rk_plan(int order, const double* alphas, const double* betas, int fold, char* out, size_t cap);   /* host-only: the per-stage buffer / term plan hb_fv_update runs for a tableau (hydro/int/rk.lua:17-44,91-165), fold != 0: with the last stage's running sum; text, one line per stage */
/* per-launch device timing of the fused stage kernel (CUDA events on the context's stream; disables graph replay
 * while enabled): total milliseconds and number of stage launches since hb_fv_profile(fv, 1) */
int hb_fv_profile(hb_fv* fv, int enable);
int hb_fv_profile_read(hb_fv* fv, double* stage_ms_total, long long* stage_launches);
/* unit-test hook: evaluate one device function per item on the GPU (kind 0 Roe flux, 1 constrainU, 2 calcDTCell,
 * 3 PLM half slope, 4 Roe flux with flux limiter); host pointers of doubles; strict selects the -fmad=false build */
int hb_debug_eval(hb_ctx* ctx, int eqn, int strict, int kind, int side, int n, const double* params, const double* aux4,
	const double* in, size_t in_count, double* out, size_t out_count);
/* host-side helper exported for tests: source index of ghost index j on an axis of ghost-inclusive size S */
int hb_ghost_source(int j, int S, int bc_min, int bc_max, int* flip_out, int* skip_out);

/* ---- multi-GPU: slab decomposition along the slowest used axis, one process per GPU.  Replaces
 *      hydro/solver/choppedup.lua:193-233,344-409 (rect copies once per step + host min of dt) by a per-stage
 *      ghost-plane exchange and a min-allreduce of dt, both over NCCL (NVLink 5 / NVSwitch). */
int hb_comm_unique_id(char* out128);                         /* ncclGetUniqueId; broadcast by the caller (e.g. torch.distributed) */
int hb_fv_comm_init(hb_fv* fv, int nranks, int rank, const char* id128);
int hb_fv_comm_destroy(hb_fv* fv);

#ifdef __cplusplus
}
#endif
#endif /* HYDROB200_H */
