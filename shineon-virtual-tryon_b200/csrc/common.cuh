// Shared helpers for libshineon_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
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

// ---------------------------------------------------------------- bf16 hi/lo planes
// x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits, so three bf16 tensor-core
// products (hi*hi + hi*lo + lo*hi) reproduce an fp32 product to ~2^-16 relative.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ float join_bf16(__nv_bfloat16 hi, __nv_bfloat16 lo) {
  return __bfloat162float(hi) + __bfloat162float(lo);
}

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
