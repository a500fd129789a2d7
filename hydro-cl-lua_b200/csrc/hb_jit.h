// hb_jit.h -- run-time compiled equation plug-ins (hb_jit.cu): the codegen seam of SURVEY 8 / north_star ("the equation-specific
// eigenvector and flux device functions are emitted by the existing Lua codegen into hand-written sm_100a CUDA kernel templates").
#pragma once
#include "hb_core.h"
#include "hb_fv_ops.h"
#include <string>

namespace hb {

struct JitProgram;     // one NVRTC-compiled module: the marching / ghost / CFL / constrain kernels instantiated over the supplied equation

// the program the FvOps of a run-time equation launch from (set by the solver around every ops call: FvOps entries carry no context)
extern thread_local JitProgram* tlsJit;

// Compiles the kernel templates over `eqnType`<real, FAST> (a class template satisfying the plug-in contract of hb_eqn_euler.cuh, defined
// by `headerSrc`, which is registered as the include file `headerName`; naming an embedded header, e.g. "hb_eqn_euler.cuh", REPLACES it).
// Returns null and fills `err` / `log` on failure.
template<class real>
JitProgram* jitCompile(hb_ctx* ctx, const char* headerName, const char* headerSrc, const char* eqnType, int dim, int slopeLimiter, bool strict,
	const double* eqnParams, std::string& err, std::string& log);
void jitFree(JitProgram* p);
template<class real> const FvOps<real>* jitOps(JitProgram* p);

}   // namespace hb
