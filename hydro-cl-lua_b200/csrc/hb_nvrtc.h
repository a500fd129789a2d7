// hb_nvrtc.h -- NVRTC and the handful of driver entry points the run-time compilation paths need (hb_module.cu, hb_jit.cu), bound at run
// time: NVRTC with dlopen, the driver through cudaGetDriverEntryPoint, so the library links neither libnvrtc nor libcuda.
#pragma once
#include "hb_core.h"
#include <dlfcn.h>
#include <cstdlib>
#include <string>
#include <vector>

namespace hb {

struct Nvrtc {
	typedef struct _nvrtcProgram* prog_t;
	int (*CreateProgram)(prog_t*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
	int (*CompileProgram)(prog_t, int, const char* const*) = nullptr;
	int (*GetCUBINSize)(prog_t, size_t*) = nullptr;
	int (*GetCUBIN)(prog_t, char*) = nullptr;
	int (*GetProgramLogSize)(prog_t, size_t*) = nullptr;
	int (*GetProgramLog)(prog_t, char*) = nullptr;
	int (*DestroyProgram)(prog_t*) = nullptr;
	int (*AddNameExpression)(prog_t, const char*) = nullptr;
	int (*GetLoweredName)(prog_t, const char*, const char**) = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
	bool ok = false;
	std::string why;
	static Nvrtc& get() {
		static Nvrtc n;
		static bool tried = false;
		if (tried) return n;
		tried = true;
		void* h = nullptr;
		std::vector<std::string> names;
		if (const char* p = getenv("HB_NVRTC_PATH")) names.push_back(p);
		names.push_back("libnvrtc.so.12");
		names.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
		names.push_back("libnvrtc.so");
		for (auto& nm : names) { h = dlopen(nm.c_str(), RTLD_NOW | RTLD_GLOBAL); if (h) break; }
		if (!h) { n.why = std::string("cannot load libnvrtc: ") + dlerror(); return n; }
#define HB_SYM(field, name) *(void**)(&n.field) = dlsym(h, name); if (!n.field) { n.why = std::string("libnvrtc lacks ") + name; return n; }
		HB_SYM(CreateProgram, "nvrtcCreateProgram") HB_SYM(CompileProgram, "nvrtcCompileProgram")
		HB_SYM(GetCUBINSize, "nvrtcGetCUBINSize") HB_SYM(GetCUBIN, "nvrtcGetCUBIN")
		HB_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize") HB_SYM(GetProgramLog, "nvrtcGetProgramLog")
		HB_SYM(DestroyProgram, "nvrtcDestroyProgram") HB_SYM(GetErrorString, "nvrtcGetErrorString")
		HB_SYM(AddNameExpression, "nvrtcAddNameExpression") HB_SYM(GetLoweredName, "nvrtcGetLoweredName")
#undef HB_SYM
		n.ok = true;
		return n;
	}
};

// the handful of driver entry points the module API needs
struct Driver {
	typedef int (*ModuleLoadData_t)(void**, const void*);
	typedef int (*ModuleUnload_t)(void*);
	typedef int (*ModuleGetFunction_t)(void**, void*, const char*);
	typedef int (*LaunchKernel_t)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**);
	typedef int (*GetErrorString_t)(int, const char**);
	typedef int (*FuncSetAttribute_t)(void*, int, int);
	ModuleLoadData_t ModuleLoadData = nullptr;
	ModuleUnload_t ModuleUnload = nullptr;
	ModuleGetFunction_t ModuleGetFunction = nullptr;
	LaunchKernel_t LaunchKernel = nullptr;
	GetErrorString_t GetErrorString = nullptr;
	FuncSetAttribute_t FuncSetAttribute = nullptr;
	bool ok = false;
	std::string why;
	static Driver& get() {
		static Driver d;
		static bool tried = false;
		if (tried) return d;
		tried = true;
		auto sym = [&](const char* name, void** out) {
			cudaDriverEntryPointQueryResult q;
			cudaError_t e = cudaGetDriverEntryPoint(name, out, cudaEnableDefault, &q);
			if (e != cudaSuccess || !*out) { d.why = std::string("driver entry point missing: ") + name; cudaGetLastError(); return false; }
			return true;
		};
		if (!sym("cuModuleLoadData", (void**)&d.ModuleLoadData)) return d;
		if (!sym("cuModuleUnload", (void**)&d.ModuleUnload)) return d;
		if (!sym("cuModuleGetFunction", (void**)&d.ModuleGetFunction)) return d;
		if (!sym("cuLaunchKernel", (void**)&d.LaunchKernel)) return d;
		if (!sym("cuGetErrorString", (void**)&d.GetErrorString)) return d;
		if (!sym("cuFuncSetAttribute", (void**)&d.FuncSetAttribute)) return d;
		d.ok = true;
		return d;
	}
	std::string err(int r) { const char* s = nullptr; if (GetErrorString) GetErrorString(r, &s); return s ? s : "unknown driver error"; }
};

}   // namespace hb
