// Resample2d forward with a shared-memory halo tile (resample2d_kernel.cu:16-72, kernel_size 1, bilinear).
//
// The per-pixel form (gather_ops.cu) issues 4 taps x C channels of 4-byte loads per output pixel, each its own 32-byte
// sector: on i.i.d. flows it is sector-bound at 0.43 of the HBM roofline (profiles/r01_memory_ops.md).  Here one CTA owns a
// 32 x 64 output tile: a single TMA box brings the source window [tile +- 12 px] of all C channels into shared memory
// (61 KB, three CTAs per SM; out-of-image box elements are zero-filled and never read, the taps are clamped to the image as
// in the reference) while the threads load their flow vectors; the taps are then shared-memory reads.  Neighbouring tiles'
// windows overlap, but those re-reads hit L2; DRAM sees the image once.  Taps outside the window (|flow| > 12 px) fall back
// to global loads, so the result is exact for any flow.
#include <cuda.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace shineon {

constexpr int kRsTW = 64, kRsTH = 32, kRsR = 12;
constexpr int kRsBW = 92;                 // >= TW + 2R + 1 = 89, rounded to a 16-byte multiple (TMA box rule)
constexpr int kRsBH = kRsTH + 2 * kRsR + 1;  // 57
constexpr int kRsPPT = kRsTW * kRsTH / 256;  // 8 output pixels per thread

template <int C>
__global__ void __launch_bounds__(256)
    resample2d_tile_kernel(const __grid_constant__ CUtensorMap tm, const float* __restrict__ in1,
                           const float* __restrict__ flow, float* __restrict__ out, int H, int W, int tiles_x) {
  pdl_grid_sync();
  extern __shared__ __align__(128) float s_tile[];  // [C][kRsBH][kRsBW]
  __shared__ __align__(8) uint64_t bar;
  const int b = blockIdx.y;
  const int x0 = ((int)blockIdx.x % tiles_x) * kRsTW, y0 = ((int)blockIdx.x / tiles_x) * kRsTH;
  const int sx0 = x0 - kRsR, sy0 = y0 - kRsR;
  const uint32_t bar_a = smem_u32(&bar);
  if (threadIdx.x == 0) {
    mbar_init(bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_a, C * kRsBH * kRsBW * 4);
    tma_load_3d(smem_u32(s_tile), &tm, bar_a, sx0, sy0, b * C);
  }
  const int HW = H * W;
  const float* fb = flow + (long)b * 2 * HW;
  const float* ib = in1 + (long)b * C * HW;
  float* ob = out + (long)b * C * HW;
  const int lx = threadIdx.x & (kRsTW - 1), ly0 = threadIdx.x / kRsTW;  // 4 rows of 64 per pass
  const int x = x0 + lx;
  float dx[kRsPPT], dy[kRsPPT];
#pragma unroll
  for (int k = 0; k < kRsPPT; ++k) {
    const int y = y0 + ly0 + 4 * k;
    const bool ok = x < W && y < H;
    dx[k] = ok ? __ldg(fb + y * W + x) : 0.f;
    dy[k] = ok ? __ldg(fb + HW + y * W + x) : 0.f;
  }
  mbar_wait(bar_a, 0);
#pragma unroll
  for (int k = 0; k < kRsPPT; ++k) {
    const int y = y0 + ly0 + 4 * k;
    if (x >= W || y >= H) continue;
    const float xf = (float)x + dx[k], yf = (float)y + dy[k];
    const float fx = floorf(xf), fy = floorf(yf);
    const float alpha = xf - fx, beta = yf - fy;  // resample2d_kernel.cu:42-43
    // int(floor(xf)) with a defined result for huge flows (same guard as the per-pixel kernel)
    const float cfx = fminf(fmaxf(fx, -4.f), (float)W + 4.f), cfy = fminf(fmaxf(fy, -4.f), (float)H + 4.f);
    const int xL = max(min((int)cfx, W - 1), 0), xR = max(min((int)cfx + 1, W - 1), 0);
    const int yT = max(min((int)cfy, H - 1), 0), yB = max(min((int)cfy + 1, H - 1), 0);
    const float w00 = (1.f - alpha) * (1.f - beta), w01 = alpha * (1.f - beta), w10 = (1.f - alpha) * beta, w11 = alpha * beta;
    const bool inside = xL >= sx0 && xR < sx0 + kRsBW && yT >= sy0 && yB < sy0 + kRsBH;
    float v[C][4];
    if (inside) {
      const int oT = (yT - sy0) * kRsBW, oB = (yB - sy0) * kRsBW, cL = xL - sx0, cR = xR - sx0;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float* pl = s_tile + c * (kRsBH * kRsBW);
        v[c][0] = pl[oT + cL]; v[c][1] = pl[oT + cR]; v[c][2] = pl[oB + cL]; v[c][3] = pl[oB + cR];
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float* pl = ib + (long)c * HW;
        v[c][0] = __ldg(pl + yT * W + xL); v[c][1] = __ldg(pl + yT * W + xR);
        v[c][2] = __ldg(pl + yB * W + xL); v[c][3] = __ldg(pl + yB * W + xR);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {  // same accumulation order as resample2d_kernel.cu:56-59 (0 + w*v ...)
      float a = 0.f;
      a += w00 * v[c][0];
      a += w01 * v[c][1];
      a += w10 * v[c][2];
      a += w11 * v[c][3];
      __stcs(ob + (long)c * HW + y * W + x, a);
    }
  }
}

}  // namespace shineon

using namespace shineon;

// Returns SHINEON_OK after launching, a positive value when the shape does not fit this variant (caller falls back).
int shineon_resample2d_tile(const float* in1, const float* flow, float* out, int B, int C, int H, int W, cudaStream_t stream) {
  if (C < 1 || C > 4 || W % 4 != 0 || (reinterpret_cast<uintptr_t>(in1) & 15) != 0 || (long)B * C > (1l << 30)) return 1;
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(SHINEON_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found (driver too old?)");
  CUtensorMap tm;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * C};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  const cuuint32_t box[3] = {kRsBW, kRsBH, (cuuint32_t)C};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(in1), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SHINEON_ERR_CUDA, "cuTensorMapEncodeTiled(resample2d) failed: CUresult %d", (int)r);
  const int smem = C * kRsBH * kRsBW * 4;
  const int tiles_x = cdiv(W, kRsTW), tiles_y = cdiv(H, kRsTH);
  const dim3 grid(tiles_x * tiles_y, B);
  cudaError_t e = cudaSuccess;
#define SHINEON_RS(CC)                                                                                              \
  {                                                                                                                 \
    static bool opted = false;                                                                                      \
    if (!opted) {                                                                                                   \
      e = cudaFuncSetAttribute(resample2d_tile_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);      \
      opted = e == cudaSuccess;                                                                                     \
    }                                                                                                               \
    if (e == cudaSuccess) klaunch(resample2d_tile_kernel<CC>, grid, 256, smem, stream, tm, in1, flow, out, H, W, tiles_x); \
  }
  switch (C) {
    case 1: SHINEON_RS(1) break;
    case 2: SHINEON_RS(2) break;
    case 3: SHINEON_RS(3) break;
    default: SHINEON_RS(4) break;
  }
#undef SHINEON_RS
  if (e != cudaSuccess) return fail(SHINEON_ERR_CUDA, "resample2d_tile: shared memory opt-in: %s", cudaGetErrorString(e));
  return after_launch("resample2d_tile_kernel");
}
