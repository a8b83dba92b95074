// Shared helpers for libshineon_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <stdlib.h>

#include <atomic>
#include <string>
#include <utility>

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

// ---------------------------------------------------------------- programmatic dependent launch
// Every kernel of the library starts with pdl_grid_sync() (griddepcontrol.wait: returns once every grid this one depends
// on has completed and flushed; a no-op for a launch without the attribute) and is launched through klaunch(), which adds
// the programmatic stream-serialisation attribute.  The next kernel of the stream (or of the captured graph) is then set
// up -- launch processing, CTA rasterisation -- while its predecessor drains, instead of after it.  Ordering is unchanged:
// nothing before the wait touches memory, and without an explicit trigger a dependent grid becomes resident only when all
// CTAs of its predecessor have exited (a chain A -> B -> C stays transitive: C passes its wait only after B completed, and B
// only after it passed its own wait on A).  Measured on B200 (profiles/r02_pdl.md): try-on step 12.03 -> 11.84 ms, FlowNet2
// 5.58 -> 5.52 ms, training step 5.47 -> 5.44 ms.  The variant that also triggers early (griddepcontrol.launch_dependents
// right after the wait, -DSHINEON_PDL_EARLY_TRIGGER) makes the successor resident next to the running persistent conv CTAs
// and was 4 % SLOWER on the try-on step.  SHINEON_PDL=0 launches with plain stream ordering.
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef SHINEON_PDL_EARLY_TRIGGER  // the measured-slower variant: successor resident while this grid still runs
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("SHINEON_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

extern thread_local cudaError_t g_launch_error;

template <typename... P, typename... A>
inline void klaunch(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  g_launch_error = cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...);
}

// Every launch goes through this: counts it and converts a launch error into a status.
inline int after_launch(const char* what) {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = g_launch_error;
  g_launch_error = cudaSuccess;
  if (e != cudaSuccess) return fail(SHINEON_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return SHINEON_OK;
}

#define SHINEON_REQUIRE(cond, ...) \
  do {                             \
    if (!(cond)) return ::shineon::fail(SHINEON_ERR_ARG, __VA_ARGS__); \
  } while (0)

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- activations
// nn.GELU() (exact erf form): 0.5 x (1 + erf(x / sqrt 2)), with erfc from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7):
//   x >= 0: x - 0.5 x erfc(z),  x < 0: 0.5 x erfc(z),  z = |x| / sqrt 2   (no cancellation in the negative tail).
// ~14 instructions with two MUFU ops; erff() costs about twice that and made the normalise + activate pass issue-bound
// (profiles/r01_memory_ops.md).  Absolute error <= 0.5 |x| 1.5e-7, four orders below the parity tolerance.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float half_erfc = 0.5f * p * t * __expf(-z * z);
  return x >= 0.f ? fmaf(-x, half_erfc, x) : x * half_erfc;
}

__device__ __forceinline__ float apply_act(float v, int act, float param) {
  switch (act) {
    case SHINEON_ACT_RELU: return fmaxf(v, 0.f);
    case SHINEON_ACT_LEAKY: return v > 0.f ? v : v * param;
    case SHINEON_ACT_GELU: return gelu_erf(v);
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
// Two values at once, as the 32-bit words the plane stores write (element `a` in the low half): one packed conversion per
// plane (F2FP.PACK_AB converts two floats per issue) and no 16-bit re-packing.  Bit-identical to split16 per element.
__device__ __forceinline__ void split16x2(float a, float b, int fmt, uint32_t& hi, uint32_t& lo) {
  if (fmt == SHINEON_FMT_FP16) {
    a = fminf(fmaxf(a, -65000.f), 65000.f);
    b = fminf(fmaxf(b, -65000.f), 65000.f);
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  } else {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
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
// Column sums of a 32 x 32 matrix held one ROW per lane (v[c] = this lane's value for column c): after five
// recursive-halving exchange steps lane l returns sum over all lanes of v[l] (31 shuffles instead of 32 x 5).
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < off) {
        const float send = up ? v[i] : v[i + off];
        const float keep = up ? v[i + off] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
  }
  return v[0];
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace shineon
