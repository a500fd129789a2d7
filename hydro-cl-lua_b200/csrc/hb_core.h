// hb_core.h -- internal definitions shared by the C-ABI translation units (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "../../include/hydrob200.h"

struct hb_ctx {
	int device = 0;
	int real_bytes = 8;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	void* cuContext = nullptr;   // primary context (driver API handle), set lazily by the module API
	double* reduceScratch = nullptr;
	int refs = 1;                // the handle itself + every live hb_buf / hb_module / hb_fv made from it
};

struct hb_buf {
	hb_ctx* ctx = nullptr;
	void* d = nullptr;
	size_t bytes = 0;
};

namespace hb {
int setError(int code, const std::string& msg);
int cudaFail(cudaError_t e, const char* what);
bool useDevice(hb_ctx* ctx);
void ctxRetain(hb_ctx* ctx);
void ctxRelease(hb_ctx* ctx);   // frees the stream/events when the last reference goes (any destruction order is safe)
#define HB_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return hb::cudaFail(e_, #expr); } while (0)
}
