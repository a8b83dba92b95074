// TMA-staged form of upconv3x3_gather (norm_act.cu has the math and the direct-from-global variant).
//
// The direct kernel is latency-bound: its ~49 dependent 16-byte gathers per thread leave only a few loads in flight
// per warp (ncu: long-scoreboard stalls 8 per issue, 25 % of HBM, profiles/r01_upconv_gather.md).  Here one elected
// thread stages the CTA's whole operand -- for each of the 9 taps a (TH+2) x (TW+2)-pixel halo tile of 32 channels --
// with nine cp.async.bulk.tensor loads (69 KB in flight per CTA, three CTAs per SM), out-of-image halo elements
// zero-filled by TMA; the 256 threads then read shared memory only.  The tap sum is also reordered: for each (fy, low-res
// row) the three fx taps are first reduced along the row with the column coefficients (12 FMAs per value instead of
// 16-20), then scattered to the two output rows.
#include "tcgen05.cuh"

namespace shineon {

constexpr int kGatherCB = 32;  // channels per CTA (one 128-byte row per pixel and tap)
constexpr int kGatherTW = 8, kGatherTH = 4;
constexpr int kGatherHaloPix = (kGatherTW + 2) * (kGatherTH + 2);
constexpr int kGatherTapFloats = kGatherHaloPix * kGatherCB;
constexpr int kGatherSmemBytes = 9 * kGatherTapFloats * 4;  // 69120

struct Up3c { float c[4][3]; };  // coefficient triples A, B, C, D (see norm_act.cu)
__device__ __forceinline__ Up3c up3c_coeffs(int i, int n) {
  Up3c u;
  const bool first = i == 0, last = i == n - 1;
  u.c[0][0] = first ? 0.f : 0.75f; u.c[0][1] = first ? 0.f : 0.25f; u.c[0][2] = 0.f;
  u.c[1][0] = first ? 0.f : 0.25f; u.c[1][1] = first ? 1.f : 0.75f; u.c[1][2] = 0.f;
  u.c[2][0] = 0.f; u.c[2][1] = last ? 1.f : 0.75f; u.c[2][2] = last ? 0.f : 0.25f;
  u.c[3][0] = 0.f; u.c[3][1] = last ? 0.f : 0.25f; u.c[3][2] = last ? 0.f : 0.75f;
  return u;
}
__host__ __device__ constexpr bool up3c_has(int f, int p, int a) { return (f + p) < 2 ? a < 2 : a > 0; }

__global__ void __launch_bounds__(256, 3)
    upconv3x3_gather_tma_kernel(const __grid_constant__ CUtensorMap tmV, const float* __restrict__ bias,
                                float* __restrict__ y, double* __restrict__ stats, int h, int w, int Cout, int tiles_x) {
  pdl_grid_sync();
  extern __shared__ __align__(128) float sv[];  // [9 taps][TH+2][TW+2][CB]
  __shared__ __align__(8) uint64_t bar;
  __shared__ float s_stat[8][kGatherCB][2];  // per warp: single writer per slot, summed in a fixed order (reproducible)
  const int n = blockIdx.z;
  const int co0 = blockIdx.y * kGatherCB;
  const int i0 = ((int)blockIdx.x / tiles_x) * kGatherTH, j0 = ((int)blockIdx.x % tiles_x) * kGatherTW;
  const uint32_t bar_a = smem_u32(&bar);
  if (threadIdx.x == 0) {
    mbar_init(bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar_a, kGatherSmemBytes);
    const uint32_t dst = smem_u32(sv);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
      tma_load_4d(dst + tap * kGatherTapFloats * 4, &tmV, bar_a, tap * Cout + co0, j0 - 1, i0 - 1, n);
  }
  __syncthreads();  // barrier initialised before anyone polls it
  const int g = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const int pi = slot / kGatherTW, pj = slot % kGatherTW;
  const int i = i0 + pi, j = j0 + pj;
  const int co = co0 + g * 4;
  const Up3c cy = up3c_coeffs(min(i, h - 1), h), cx = up3c_coeffs(min(j, w - 1), w);
  float out[2][2][4];
  {
    float4 b4 = bias ? __ldg(reinterpret_cast<const float4*>(bias + co)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = 0; q < 2; ++q) { out[p][q][0] = b4.x; out[p][q][1] = b4.y; out[p][q][2] = b4.z; out[p][q][3] = b4.w; }
  }
  mbar_wait(bar_a, 0);
  // halo tile origin is (i0-1, j0-1): neighbour a (0,1,2 = i-1,i,i+1) of this pixel sits at halo row pi + a
  const float* base = sv + (pi * (kGatherTW + 2) + pj) * kGatherCB + g * 4;
#pragma unroll
  for (int fy = 0; fy < 3; ++fy) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (!(up3c_has(fy, 0, a) || up3c_has(fy, 1, a))) continue;
      float cs[2][4];  // sum over fx of the column-interpolated row, for the two output columns
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) cs[q][k] = 0.f;
#pragma unroll
      for (int fx = 0; fx < 3; ++fx) {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          if (!(up3c_has(fx, 0, b) || up3c_has(fx, 1, b))) continue;
          const float4 v = *reinterpret_cast<const float4*>(base + (fy * 3 + fx) * kGatherTapFloats +
                                                            (a * (kGatherTW + 2) + b) * kGatherCB);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (!up3c_has(fx, q, b)) continue;
            const float wx = cx.c[fx + q][b];
            cs[q][0] = fmaf(wx, v.x, cs[q][0]);
            cs[q][1] = fmaf(wx, v.y, cs[q][1]);
            cs[q][2] = fmaf(wx, v.z, cs[q][2]);
            cs[q][3] = fmaf(wx, v.w, cs[q][3]);
          }
        }
      }
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        if (!up3c_has(fy, p, a)) continue;
        const float wy = cy.c[fy + p][a];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int k = 0; k < 4; ++k) out[p][q][k] = fmaf(wy, cs[q][k], out[p][q][k]);
      }
    }
  }
  const bool ok = i < h && j < w;
  if (ok) {
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float* dst = y + (((long)n * 2 * h + 2 * i + p) * (2 * w) + 2 * j + q) * Cout + co;
        *reinterpret_cast<float4*>(dst) = make_float4(out[p][q][0], out[p][q][1], out[p][q][2], out[p][q][3]);
      }
  }
  if (stats != nullptr) {
    // InstanceNorm statistics of the f32 output (the whole CTA lies in image n): per-thread sums over its 2x2 pixels,
    // lanes sharing a channel group (g = lane & 7) combined by shuffles, warps through shared atomics, one pair of
    // f64 atomics per channel and CTA
    float s4[4], q4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float sacc = 0.f, qacc = 0.f;
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float x = ok ? out[p][q][k] : 0.f;
          sacc += x;
          qacc = fmaf(x, x, qacc);
        }
      sacc += __shfl_xor_sync(0xffffffffu, sacc, 8);
      qacc += __shfl_xor_sync(0xffffffffu, qacc, 8);
      sacc += __shfl_xor_sync(0xffffffffu, sacc, 16);
      qacc += __shfl_xor_sync(0xffffffffu, qacc, 16);
      s4[k] = sacc;
      q4[k] = qacc;
    }
    if ((threadIdx.x & 31) < 8) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s_stat[threadIdx.x >> 5][g * 4 + k][0] = s4[k];
        s_stat[threadIdx.x >> 5][g * 4 + k][1] = q4[k];
      }
    }
    __syncthreads();
    if (threadIdx.x < 2 * kGatherCB) {
      const int c = threadIdx.x >> 1, which = threadIdx.x & 1;
      double tot = 0.0;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) tot += (double)s_stat[wv][c][which];
      atomicAdd(stats + ((long)n * Cout + co0 + c) * 2 + which, tot);
    }
  }
}

}  // namespace shineon

using namespace shineon;

// Returns SHINEON_OK after launching, or a positive value when the shape does not fit this variant (caller falls back).
int shineon_upconv3x3_gather_tma(const float* t, const float* bias, float* y, double* stats_ws, int N, int h, int w, int Cout,
                                 int tstride, cudaStream_t stream) {
  if (Cout % kGatherCB != 0 || tstride % 4 != 0 || (reinterpret_cast<uintptr_t>(t) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(y) & 15) != 0 || (bias && (reinterpret_cast<uintptr_t>(bias) & 15) != 0))
    return 1;
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(SHINEON_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found (driver too old?)");
  CUtensorMap tm;
  const cuuint64_t dims[4] = {(cuuint64_t)tstride, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)tstride * 4, (cuuint64_t)w * tstride * 4, (cuuint64_t)h * w * tstride * 4};
  const cuuint32_t box[4] = {kGatherCB, kGatherTW + 2, kGatherTH + 2, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(t), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SHINEON_ERR_CUDA, "cuTensorMapEncodeTiled(upconv gather) failed: CUresult %d", (int)r);
  static bool opted = false;
  if (!opted) {
    cudaError_t e = cudaFuncSetAttribute(upconv3x3_gather_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGatherSmemBytes);
    if (e != cudaSuccess) return fail(SHINEON_ERR_CUDA, "upconv3x3_gather: shared memory opt-in: %s", cudaGetErrorString(e));
    opted = true;
  }
  const int tiles_x = cdiv(w, kGatherTW), tiles_y = cdiv(h, kGatherTH);
  dim3 grid(tiles_x * tiles_y, Cout / kGatherCB, N);
  klaunch(upconv3x3_gather_tma_kernel, grid, 256, kGatherSmemBytes, stream, tm, bias, y, stats_ws, h, w, Cout, tiles_x);
  return after_launch("upconv3x3_gather_tma_kernel");
}
