// hb_rtc_compat.h -- lets the kernel headers (hb_math.cuh ... hb_fv_march3.cuh) compile under NVRTC, which has no host standard library:
// the fixed-width integer types, HUGE_VAL and the opaque TMA descriptor type are provided here; under nvcc / g++ the real headers are used.
#pragma once
#if defined(__CUDACC_RTC__)
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#ifndef HUGE_VAL
#define HUGE_VAL (__longlong_as_double(0x7ff0000000000000LL))
#endif
#ifndef HUGE_VALF
#define HUGE_VALF (__int_as_float(0x7f800000))
#endif
// cuda.h: the 128-byte tensor-map descriptor cuTensorMapEncodeTiled fills on the host; opaque to device code
struct alignas(64) CUtensorMap_st { unsigned long long opaque[16]; };
typedef CUtensorMap_st CUtensorMap;
#else
#include <cmath>
#include <cstdint>
#include <cstring>
#if defined(__CUDACC__)
#include <cuda.h>
#endif
#endif
