// Shared device/host helpers for the hiarch_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
#include "../../include/hiarch_b200.h"

#define HIA_API extern "C" __attribute__((visibility("default")))

namespace hia {

void set_last_cuda_error(cudaError_t e);
void count_launch();          // per-process counter of kernels launched by this library
int sel_aggregate();          // 1 (default): warp-aggregated histogram adds (match.any); HIA_SEL_AGG=0 -> plain atomics

inline int cuda_ok(cudaError_t e) {
  if (e == cudaSuccess) return HIA_OK;
  set_last_cuda_error(e);
  return HIA_ERR_CUDA;
}

#define HIA_CUDA(call)                                   \
  do {                                                   \
    int _s = ::hia::cuda_ok((call));                     \
    if (_s != HIA_OK) return _s;                         \
  } while (0)

#define HIA_LAUNCH_CHECK()                 \
  do {                                     \
    ::hia::count_launch();                 \
    HIA_CUDA(cudaGetLastError());          \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Streaming multiprocessors of the current device (148 on B200), queried once per process.
int num_sms();

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Positive floats (and +inf) order like their bit patterns.
__device__ __forceinline__ void atomic_min_pos_float(float* addr, float v) {
  atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
// Order-preserving float <-> uint32 key (any sign): smaller float <=> smaller key.
__host__ __device__ __forceinline__ unsigned int float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
  unsigned int b = __float_as_uint(f);
#else
  unsigned int b; memcpy(&b, &f, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(unsigned int k) {
  unsigned int b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f; memcpy(&f, &b, 4); return f;
#endif
}

// streaming 128-bit accesses (no L1 allocation)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// L2 residency hints for two-pass kernels: the first pass asks L2 to KEEP the line (evict_last), the
// second (last) use and the output stream are evict-first, so that the outputs do not push the
// lines of rows still waiting for their second pass out of the 126 MB L2.
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ld_keep4(const float* p, unsigned long long pol) {
  float4 r;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ float4 ld_last4(const float* p) {
  float4 r;
  asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_cs_u2(void* p, const uint2& v) {
  asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_cs_f4(void* p, const float4& v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace hia
