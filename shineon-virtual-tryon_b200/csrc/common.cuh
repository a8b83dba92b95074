// Shared helpers for libshineon_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/shineon_b200.h"

namespace shineon {

extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launch_count;

inline int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

// Every launch goes through this: counts it and converts a launch error into a status.
inline int after_launch(const char* what) {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SHINEON_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return SHINEON_OK;
}

#define SHINEON_REQUIRE(cond, ...) \
  do {                             \
    if (!(cond)) return ::shineon::fail(SHINEON_ERR_ARG, __VA_ARGS__); \
  } while (0)

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- activations
__device__ __forceinline__ float apply_act(float v, int act, float param) {
  switch (act) {
    case SHINEON_ACT_RELU: return fmaxf(v, 0.f);
    case SHINEON_ACT_LEAKY: return v > 0.f ? v : v * param;
    case SHINEON_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    case SHINEON_ACT_SWISH: return v / (1.f + expf(-v));
    case SHINEON_ACT_SINE: return sinf(30.f * v);
    case SHINEON_ACT_TANH: return tanhf(v);
    case SHINEON_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

// ---------------------------------------------------------------- 16-bit hi/lo planes
// x ~= hi + lo with hi = round16(x), lo = round16(x - hi).  Three tensor-core products
// (hi*hi + hi*lo + lo*hi) then reproduce an fp32 product to ~2^-16 (bf16: 8+8 mantissa bits) or
// ~2^-21 (fp16: 11+11 bits).  Planes are stored as raw 16-bit words; `fmt` selects the encoding.
typedef uint16_t plane_t;
__device__ __forceinline__ void split16(float x, int fmt, plane_t& hi, plane_t& lo) {
  if (fmt == SHINEON_FMT_FP16) {
    x = fminf(fmaxf(x, -65000.f), 65000.f);  // fp16 range guard (normalised activations never get close)
    const __half h = __float2half_rn(x);
    const __half l = __float2half_rn(x - __half2float(h));
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(l);
  }
}
__device__ __forceinline__ float load16(plane_t v, int fmt) {
  return fmt == SHINEON_FMT_FP16 ? __half2float(__ushort_as_half(v)) : __bfloat162float(__ushort_as_bfloat16(v));
}
__device__ __forceinline__ float join16(plane_t hi, plane_t lo, int fmt) { return load16(hi, fmt) + load16(lo, fmt); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace shineon
