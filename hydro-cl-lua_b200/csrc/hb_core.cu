// hb_core.cu -- contexts, buffers, timers and reductions of the C ABI (include/hydrob200.h).
// Replaces the lua-opencl objects the reference's hot path uses: CLEnv, CLBuffer, env:reduce.
#include "hb_core.h"
#include <cstdio>
#include <cstring>
#include <cmath>

namespace hb {

static thread_local std::string g_lastError;

int setError(int code, const std::string& msg) { g_lastError = msg; return code; }
int cudaFail(cudaError_t e, const char* what) {
	int code = (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? HB_ERR_NO_DEVICE : HB_ERR_CUDA;
	// clear the sticky-less error so later calls report their own failure
	cudaGetLastError();
	return setError(code, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
bool useDevice(hb_ctx* ctx) { return cudaSetDevice(ctx->device) == cudaSuccess; }
void ctxRetain(hb_ctx* c) { if (c) c->refs++; }
void ctxRelease(hb_ctx* c) {
	if (!c || --c->refs > 0) return;
	useDevice(c);
	cudaStreamSynchronize(c->stream);
	if (c->reduceScratch) cudaFree(c->reduceScratch);
	if (c->ev0) cudaEventDestroy(c->ev0);
	if (c->ev1) cudaEventDestroy(c->ev1);
	if (c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

// element-generic grid-stride reduction into a double scalar (ordered-bits atomics for min/max, atomicAdd for sum)
template<class real, int OP>
__global__ void reduce_kernel(const real* __restrict__ x, size_t n, double* out, unsigned long long* outBits)
{
	__shared__ double red[32];
	double v = OP == HB_REDUCE_MIN ? HUGE_VAL : (OP == HB_REDUCE_MAX ? -HUGE_VAL : 0.);
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		double const u = double(x[i]);
		if (OP == HB_REDUCE_MIN) v = u < v ? u : v;
		else if (OP == HB_REDUCE_MAX) v = v < u ? u : v;
		else v += u;
	}
	for (int o = 16; o > 0; o >>= 1) {
		double const u = __shfl_xor_sync(0xffffffffu, v, o);
		if (OP == HB_REDUCE_MIN) v = u < v ? u : v;
		else if (OP == HB_REDUCE_MAX) v = v < u ? u : v;
		else v += u;
	}
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
	__syncthreads();
	if (threadIdx.x == 0) {
		int const nw = (blockDim.x + 31) / 32;
		for (int w = 1; w < nw; ++w) {
			double const u = red[w];
			if (OP == HB_REDUCE_MIN) v = u < v ? u : v;
			else if (OP == HB_REDUCE_MAX) v = v < u ? u : v;
			else v += u;
		}
		if (OP == HB_REDUCE_SUM) atomicAdd(out, v);
		else {
			// order-preserving map of a double onto unsigned 64-bit
			long long b = __double_as_longlong(v);
			unsigned long long key = b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
			if (OP == HB_REDUCE_MIN) atomicMin(outBits, key); else atomicMax(outBits, key);
		}
	}
}

}   // namespace hb

using namespace hb;

extern "C" {

const char* hb_last_error(void) { return g_lastError.c_str(); }
int hb_version(void) { return 100; }

int hb_device_count(int* count) {
	if (!count) return setError(HB_ERR_INVALID, "hb_device_count: null argument");
	*count = 0;
	cudaError_t e = cudaGetDeviceCount(count);
	if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) { cudaGetLastError(); *count = 0; return HB_OK; }   // no driver / no device: count 0
	if (e != cudaSuccess) return cudaFail(e, "cudaGetDeviceCount");
	return HB_OK;
}

int hb_ctx_create(int device, int real_bytes, hb_ctx** out) {
	if (!out) return setError(HB_ERR_INVALID, "hb_ctx_create: null out");
	*out = nullptr;
	if (real_bytes != 8 && real_bytes != 4) return setError(HB_ERR_INVALID, "hb_ctx_create: real_bytes must be 8 or 4");
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || (e == cudaSuccess && n == 0)) {
		cudaGetLastError();
		return setError(HB_ERR_NO_DEVICE, std::string("hb_ctx_create: no usable CUDA device (") + (e == cudaSuccess ? "device count 0" : cudaGetErrorName(e))
			+ "); there is no CPU fallback");
	}
	if (e != cudaSuccess) return cudaFail(e, "cudaGetDeviceCount");
	if (device < 0 || device >= n) return setError(HB_ERR_INVALID, "hb_ctx_create: device index out of range");
	HB_CUDA(cudaSetDevice(device));
	hb_ctx* c = new hb_ctx();
	c->device = device; c->real_bytes = real_bytes;
	e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
	if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
	if (e == cudaSuccess) e = cudaMalloc(&c->reduceScratch, 64);
	if (e != cudaSuccess) { delete c; return cudaFail(e, "hb_ctx_create"); }
	*out = c;
	return HB_OK;
}

int hb_ctx_destroy(hb_ctx* c) {
	if (!c) return HB_OK;
	useDevice(c);
	cudaStreamSynchronize(c->stream);
	hb::ctxRelease(c);
	return HB_OK;
}

int hb_ctx_real_bytes(hb_ctx* c) { return c ? c->real_bytes : 0; }

int hb_device_name(hb_ctx* c, char* out, size_t cap) {
	if (!c || !out || !cap) return setError(HB_ERR_INVALID, "hb_device_name: bad argument");
	cudaDeviceProp p;
	HB_CUDA(cudaGetDeviceProperties(&p, c->device));
	snprintf(out, cap, "%s", p.name);
	return HB_OK;
}
int hb_device_max_threads(hb_ctx* c, int* out) {
	if (!c || !out) return setError(HB_ERR_INVALID, "hb_device_max_threads: bad argument");
	HB_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMaxThreadsPerBlock, c->device));
	return HB_OK;
}
int hb_device_sm_count(hb_ctx* c, int* out) {
	if (!c || !out) return setError(HB_ERR_INVALID, "hb_device_sm_count: bad argument");
	HB_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, c->device));
	return HB_OK;
}
int hb_sync(hb_ctx* c) {
	if (!c) return setError(HB_ERR_INVALID, "hb_sync: null ctx");
	useDevice(c);
	HB_CUDA(cudaStreamSynchronize(c->stream));
	return HB_OK;
}
void* hb_ctx_stream(hb_ctx* c) { return c ? (void*)c->stream : nullptr; }

int hb_timer_start(hb_ctx* c) {
	if (!c) return setError(HB_ERR_INVALID, "hb_timer_start: null ctx");
	useDevice(c);
	HB_CUDA(cudaEventRecord(c->ev0, c->stream));
	return HB_OK;
}
int hb_timer_stop(hb_ctx* c, float* ms) {
	if (!c || !ms) return setError(HB_ERR_INVALID, "hb_timer_stop: bad argument");
	useDevice(c);
	HB_CUDA(cudaEventRecord(c->ev1, c->stream));
	HB_CUDA(cudaEventSynchronize(c->ev1));
	HB_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
	return HB_OK;
}

int hb_host_alloc(size_t bytes, void** out) {
	if (!out) return setError(HB_ERR_INVALID, "hb_host_alloc: null out");
	HB_CUDA(cudaMallocHost(out, bytes));
	return HB_OK;
}
int hb_host_free(void* p) {
	if (p) HB_CUDA(cudaFreeHost(p));
	return HB_OK;
}

// ---- buffers
int hb_buf_alloc(hb_ctx* c, size_t bytes, hb_buf** out) {
	if (!c || !out) return setError(HB_ERR_INVALID, "hb_buf_alloc: bad argument");
	*out = nullptr;
	useDevice(c);
	void* d = nullptr;
	HB_CUDA(cudaMalloc(&d, bytes ? bytes : 1));
	hb_buf* b = new hb_buf();
	b->ctx = c; b->d = d; b->bytes = bytes;
	hb::ctxRetain(c);
	*out = b;
	return HB_OK;
}
int hb_buf_free(hb_buf* b) {
	if (!b) return HB_OK;
	useDevice(b->ctx);
	cudaStreamSynchronize(b->ctx->stream);
	cudaFree(b->d);
	hb::ctxRelease(b->ctx);
	delete b;
	return HB_OK;
}
size_t hb_buf_size(hb_buf* b) { return b ? b->bytes : 0; }
void* hb_buf_devptr(hb_buf* b) { return b ? b->d : nullptr; }

static int checkRange(hb_buf* b, size_t off, size_t n, const char* who) {
	if (!b) return setError(HB_ERR_INVALID, std::string(who) + ": null buffer");
	if (off > b->bytes || n > b->bytes - off) return setError(HB_ERR_INVALID, std::string(who) + ": range outside the buffer");
	return HB_OK;
}
int hb_buf_write(hb_buf* b, const void* host, size_t off, size_t n) {
	if (int r = checkRange(b, off, n, "hb_buf_write")) return r;
	if (!host && n) return setError(HB_ERR_INVALID, "hb_buf_write: null host pointer");
	useDevice(b->ctx);
	HB_CUDA(cudaMemcpyAsync((char*)b->d + off, host, n, cudaMemcpyHostToDevice, b->ctx->stream));
	return HB_OK;
}
int hb_buf_read(hb_buf* b, void* host, size_t off, size_t n) {
	if (int r = checkRange(b, off, n, "hb_buf_read")) return r;
	if (!host && n) return setError(HB_ERR_INVALID, "hb_buf_read: null host pointer");
	useDevice(b->ctx);
	HB_CUDA(cudaMemcpyAsync(host, (char*)b->d + off, n, cudaMemcpyDeviceToHost, b->ctx->stream));
	HB_CUDA(cudaStreamSynchronize(b->ctx->stream));
	return HB_OK;
}

namespace hb {
__global__ void fill_pattern(unsigned char* dst, size_t n, size_t pb, unsigned long long p0, unsigned long long p1) {
	unsigned long long pat[2] = {p0, p1};
	unsigned char const* pc = reinterpret_cast<unsigned char const*>(pat);
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = pc[i % pb];
}
}
int hb_buf_fill(hb_buf* b, const void* pattern, size_t pb, size_t off, size_t n) {
	if (int r = checkRange(b, off, n, "hb_buf_fill")) return r;
	if (!pattern || pb == 0 || pb > 16 || n % pb) return setError(HB_ERR_INVALID, "hb_buf_fill: pattern must be 1..16 bytes and divide the range");
	useDevice(b->ctx);
	bool allSame = true;
	for (size_t i = 1; i < pb; ++i) allSame = allSame && ((const unsigned char*)pattern)[i] == ((const unsigned char*)pattern)[0];
	if (allSame) {
		HB_CUDA(cudaMemsetAsync((char*)b->d + off, ((const unsigned char*)pattern)[0], n, b->ctx->stream));
		return HB_OK;
	}
	unsigned long long p[2] = {0, 0};
	memcpy(p, pattern, pb);
	size_t blocks = (n + 255) / 256; if (blocks > 148 * 32) blocks = 148 * 32;
	hb::fill_pattern<<<(unsigned)blocks, 256, 0, b->ctx->stream>>>((unsigned char*)b->d + off, n, pb, p[0], p[1]);
	HB_CUDA(cudaGetLastError());
	return HB_OK;
}
int hb_buf_copy(hb_buf* dst, size_t doff, hb_buf* src, size_t soff, size_t n) {
	if (int r = checkRange(dst, doff, n, "hb_buf_copy(dst)")) return r;
	if (int r = checkRange(src, soff, n, "hb_buf_copy(src)")) return r;
	useDevice(dst->ctx);
	HB_CUDA(cudaMemcpyAsync((char*)dst->d + doff, (const char*)src->d + soff, n, cudaMemcpyDeviceToDevice, dst->ctx->stream));
	return HB_OK;
}
int hb_buf_copy_rect(hb_buf* dst, hb_buf* src, const size_t so[3], const size_t dor[3], const size_t region[3],
	size_t srp, size_t ssp, size_t drp, size_t dsp)
{
	if (!dst || !src || !so || !dor || !region) return setError(HB_ERR_INVALID, "hb_buf_copy_rect: null argument");
	if (srp == 0) srp = region[0];
	if (drp == 0) drp = region[0];
	if (ssp == 0) ssp = srp * region[1];
	if (dsp == 0) dsp = drp * region[1];
	if (region[0] == 0 || region[1] == 0 || region[2] == 0) return HB_OK;
	size_t const sEnd = (so[2] + region[2] - 1) * ssp + (so[1] + region[1] - 1) * srp + so[0] + region[0];
	size_t const dEnd = (dor[2] + region[2] - 1) * dsp + (dor[1] + region[1] - 1) * drp + dor[0] + region[0];
	if (sEnd > src->bytes || dEnd > dst->bytes) return setError(HB_ERR_INVALID, "hb_buf_copy_rect: region outside a buffer");
	useDevice(dst->ctx);
	for (size_t z = 0; z < region[2]; ++z) {
		const char* s = (const char*)src->d + (so[2] + z) * ssp + so[1] * srp + so[0];
		char* d = (char*)dst->d + (dor[2] + z) * dsp + dor[1] * drp + dor[0];
		HB_CUDA(cudaMemcpy2DAsync(d, drp, s, srp, region[0], region[1], cudaMemcpyDeviceToDevice, dst->ctx->stream));
	}
	return HB_OK;
}

int hb_reduce(hb_ctx* c, hb_buf* b, size_t count, int op, double* host_out) {
	if (!c || !b || !host_out) return setError(HB_ERR_INVALID, "hb_reduce: null argument");
	if (op < 0 || op > 2) return setError(HB_ERR_INVALID, "hb_reduce: unknown op");
	if (count * (size_t)c->real_bytes > b->bytes) return setError(HB_ERR_INVALID, "hb_reduce: count exceeds the buffer");
	useDevice(c);
	double init = 0; unsigned long long initBits = op == HB_REDUCE_MIN ? ~0ull : 0ull;
	double* dOut = c->reduceScratch;
	unsigned long long* dBits = reinterpret_cast<unsigned long long*>(c->reduceScratch + 1);
	HB_CUDA(cudaMemcpyAsync(dOut, &init, 8, cudaMemcpyHostToDevice, c->stream));
	HB_CUDA(cudaMemcpyAsync(dBits, &initBits, 8, cudaMemcpyHostToDevice, c->stream));
	size_t blocks = (count + 255) / 256; if (blocks > 148 * 8) blocks = 148 * 8; if (blocks == 0) blocks = 1;
#define HB_RED(T, OP) hb::reduce_kernel<T, OP><<<(unsigned)blocks, 256, 0, c->stream>>>((const T*)b->d, count, dOut, dBits)
	if (c->real_bytes == 8) { if (op == 0) HB_RED(double, 0); else if (op == 1) HB_RED(double, 1); else HB_RED(double, 2); }
	else { if (op == 0) HB_RED(float, 0); else if (op == 1) HB_RED(float, 1); else HB_RED(float, 2); }
#undef HB_RED
	HB_CUDA(cudaGetLastError());
	double sum; unsigned long long key;
	HB_CUDA(cudaMemcpyAsync(&sum, dOut, 8, cudaMemcpyDeviceToHost, c->stream));
	HB_CUDA(cudaMemcpyAsync(&key, dBits, 8, cudaMemcpyDeviceToHost, c->stream));
	HB_CUDA(cudaStreamSynchronize(c->stream));
	if (op == HB_REDUCE_SUM) *host_out = sum;
	else if (count == 0) *host_out = op == HB_REDUCE_MIN ? HUGE_VAL : -HUGE_VAL;
	else {
		unsigned long long bts = (key & 0x8000000000000000ull) ? (key & 0x7fffffffffffffffull) : ~key;
		memcpy(host_out, &bts, 8);
	}
	return HB_OK;
}

}   // extern "C"
