// hb_module.cu -- run-time compiled programs and generic kernel launch of the C ABI (hb_module_*, hb_kernel_*).
//
// Replaces lua-opencl's Program / kernel objects as the reference uses them:
//   Program{name,code}:compile{buildOptions} + cache/<ident>/bin  hydro/solver/solverbase.lua:558-713,1696-1699
//   program:kernel(name, args...) ; k.obj:setArg(i, x) ; k(...)     hydro/solver/fvsolver.lua:216-221,
//                                                                   hydro/solver/solverbase.lua:1328-1345
//   cmds:enqueueNDRangeKernel{kernel, globalSize, localSize}        hydro/solver/gridsolver.lua:1272-1314
// Source is CUDA C++; NVRTC compiles it straight to an sm_100a cubin, which the driver loads into the context's
// primary CUDA context.  NVRTC is bound with dlopen and the driver entry points come from
// cudaGetDriverEntryPoint, so the library itself links neither libnvrtc nor libcuda.
// Optional on-disk cubin cache: directory in $HB_CACHE_DIR, keyed by a hash of source + options.
#include "hb_core.h"
#include "hb_nvrtc.h"
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

namespace hb {

static unsigned long long fnv1a(const std::string& s, unsigned long long h = 1469598103934665603ull) {
	for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
	return h;
}

}   // namespace hb

struct hb_module {
	hb_ctx* ctx = nullptr;
	void* cuModule = nullptr;
	std::string name;
};
struct hb_kernel {
	hb_module* mod = nullptr;
	void* cuFunction = nullptr;
	std::string name;
	std::vector<std::vector<unsigned char>> args;
	std::vector<bool> isSet;
};

using namespace hb;

extern "C" {

int hb_module_compile(hb_ctx* ctx, const char* src, const char* name, const char* const* opts, int nopts,
	hb_module** out, char* log, size_t logCap)
{
	if (log && logCap) log[0] = 0;
	if (!ctx || !src || !out) return setError(HB_ERR_INVALID, "hb_module_compile: null argument");
	*out = nullptr;
	if (!useDevice(ctx)) return setError(HB_ERR_NO_DEVICE, "hb_module_compile: cannot select the device");
	HB_CUDA(cudaFree(0));   // make sure the primary context exists and is current
	Driver& D = Driver::get();
	if (!D.ok) return setError(HB_ERR_CUDA, "hb_module_compile: " + D.why);

	std::vector<const char*> o;
	std::string key = src;
	o.push_back("--gpu-architecture=sm_100a");
	bool userStd = false;
	for (int i = 0; i < nopts; ++i) if (opts && opts[i] && (!strncmp(opts[i], "--std", 5) || !strncmp(opts[i], "-std", 4))) userStd = true;
	if (!userStd) o.push_back("--std=c++17");
	// the solver's floating-point type (hydro/app.lua:892,926-929).  A macro named `real` would also rewrite every `template<class real>` of
	// an included header, so CUDA source gets HB_REAL and writes its own `typedef HB_REAL real;` (hb_module_compile_opencl prepends it)
	o.push_back(ctx->real_bytes == 8 ? "-DHB_REAL=double" : "-DHB_REAL=float");
	for (int i = 0; i < nopts; ++i) if (opts && opts[i]) o.push_back(opts[i]);
	for (auto s : o) { key += '\n'; key += s; }

	std::vector<char> cubin;
	std::string cachePath;
	if (const char* dir = getenv("HB_CACHE_DIR")) {
		char buf[64]; snprintf(buf, sizeof(buf), "%016llx", fnv1a(key));
		cachePath = std::string(dir) + "/" + (name ? name : "module") + "-" + buf + ".cubin";
		std::ifstream f(cachePath, std::ios::binary);
		if (f) cubin.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
	}
	if (cubin.empty()) {
		Nvrtc& N = Nvrtc::get();
		if (!N.ok) return setError(HB_ERR_CUDA, "hb_module_compile: " + N.why);
		Nvrtc::prog_t prog = nullptr;
		int r = N.CreateProgram(&prog, src, name ? name : "module.cu", 0, nullptr, nullptr);
		if (r) return setError(HB_ERR_CUDA, std::string("nvrtcCreateProgram: ") + N.GetErrorString(r));
		r = N.CompileProgram(prog, (int)o.size(), o.data());
		size_t ls = 0;
		N.GetProgramLogSize(prog, &ls);
		std::string lg(ls, 0);
		if (ls) N.GetProgramLog(prog, &lg[0]);
		if (log && logCap) snprintf(log, logCap, "%s", lg.c_str());
		if (r) {
			N.DestroyProgram(&prog);
			return setError(HB_ERR_COMPILE, std::string("nvrtcCompileProgram: ") + N.GetErrorString(r) + "\n" + lg);
		}
		size_t cs = 0;
		r = N.GetCUBINSize(prog, &cs);
		if (!r) { cubin.resize(cs); r = N.GetCUBIN(prog, cubin.data()); }
		N.DestroyProgram(&prog);
		if (r) return setError(HB_ERR_CUDA, std::string("nvrtcGetCUBIN: ") + N.GetErrorString(r));
		if (!cachePath.empty()) { std::ofstream f(cachePath, std::ios::binary); f.write(cubin.data(), (std::streamsize)cubin.size()); }
	}
	void* mod = nullptr;
	int r = D.ModuleLoadData(&mod, cubin.data());
	if (r) return setError(HB_ERR_CUDA, "cuModuleLoadData: " + D.err(r));
	hb_module* m = new hb_module();
	m->ctx = ctx; m->cuModule = mod; m->name = name ? name : "module";
	ctxRetain(ctx);
	*out = m;
	return HB_OK;
}

int hb_module_free(hb_module* m) {
	if (!m) return HB_OK;
	useDevice(m->ctx);
	cudaStreamSynchronize(m->ctx->stream);
	Driver::get().ModuleUnload(m->cuModule);
	ctxRelease(m->ctx);
	delete m;
	return HB_OK;
}

int hb_kernel_get(hb_module* m, const char* name, hb_kernel** out) {
	if (!m || !name || !out) return setError(HB_ERR_INVALID, "hb_kernel_get: null argument");
	*out = nullptr;
	useDevice(m->ctx);
	Driver& D = Driver::get();
	void* f = nullptr;
	int r = D.ModuleGetFunction(&f, m->cuModule, name);
	if (r) return setError(HB_ERR_INVALID, std::string("hb_kernel_get: no kernel '") + name + "' in module " + m->name + ": " + D.err(r));
	hb_kernel* k = new hb_kernel();
	k->mod = m; k->cuFunction = f; k->name = name;
	*out = k;
	return HB_OK;
}

// kernel objects are owned by the caller (program:kernel(...) wrappers are garbage-collected Lua objects in the reference); a kernel must
// be freed before its module
int hb_kernel_free(hb_kernel* k) {
	delete k;
	return HB_OK;
}

static int setArg(hb_kernel* k, int index, const void* v, size_t bytes) {
	if (!k || index < 0 || index > 63 || (!v && bytes) || bytes > 4096) return setError(HB_ERR_INVALID, "hb_kernel_set_arg: bad argument");
	if ((int)k->args.size() <= index) { k->args.resize(index + 1); k->isSet.resize(index + 1, false); }
	k->args[index].assign((const unsigned char*)v, (const unsigned char*)v + bytes);
	k->isSet[index] = true;
	return HB_OK;
}
int hb_kernel_set_arg(hb_kernel* k, int index, const void* value, size_t bytes) { return setArg(k, index, value, bytes); }
int hb_kernel_set_arg_buf(hb_kernel* k, int index, hb_buf* buf) {
	void* p = buf ? buf->d : nullptr;
	return setArg(k, index, &p, sizeof(p));
}

int hb_kernel_launch(hb_kernel* k, const size_t gs[3], const size_t ls[3], size_t smem) {
	if (!k || !gs || !ls) return setError(HB_ERR_INVALID, "hb_kernel_launch: null argument");
	hb_ctx* ctx = k->mod->ctx;
	useDevice(ctx);
	Driver& D = Driver::get();
	std::vector<void*> argv(k->args.size());
	for (size_t i = 0; i < k->args.size(); ++i) {
		if (!k->isSet[i]) return setError(HB_ERR_INVALID, "hb_kernel_launch: argument " + std::to_string(i) + " of " + k->name + " was never set");
		argv[i] = k->args[i].data();
	}
	unsigned grid[3], block[3];
	for (int a = 0; a < 3; ++a) {
		size_t const l = ls[a] ? ls[a] : 1, g = gs[a] ? gs[a] : 1;
		block[a] = (unsigned)l;
		grid[a] = (unsigned)((g + l - 1) / l);        // OpenCL semantics: global size counts work-items
	}
	if (smem > 48 * 1024) {
		int r = D.FuncSetAttribute(k->cuFunction, 8 /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES */, (int)smem);
		if (r) return setError(HB_ERR_CUDA, "cuFuncSetAttribute: " + D.err(r));
	}
	int r = D.LaunchKernel(k->cuFunction, grid[0], grid[1], grid[2], block[0], block[1], block[2], (unsigned)smem, (void*)ctx->stream,
		argv.empty() ? nullptr : argv.data(), nullptr);
	if (r) return setError(HB_ERR_CUDA, "cuLaunchKernel(" + k->name + "): " + D.err(r));
	return HB_OK;
}

}   // extern "C"
