// Gather-type kernels of the try-on hot path (HBM/L2-bound, fp32):
//   * TPS grid generation          — reference: models/networks/cpvton/warp.py:191-318
//   * bilinear grid_sample         — reference: F.grid_sample call sites models/warp_model.py:85-86,143-145
//   * fused TPS + grid_sample      — the grid is never written to HBM
//   * Resample2d fwd/bwd           — reference: resample2d_kernel.cu:16-72, :76-125, :128-198
//   * ChannelNorm fwd/bwd          — reference: channelnorm_kernel.cu:19-60, :64-96
// All tensors here are the reference's public NCHW fp32 layout.
#include <stdarg.h>

#include "common.cuh"

namespace shineon {

constexpr int kMaxTpsN = 64;  // grid_size <= 8

// ---------------------------------------------------------------------------------------------
// TPS coefficients: W = Li[:N,:N] Q, A = Li[N:,:N] Q with Q = theta + P_base  (warp.py:207-249)
// Computed redundantly by every CTA (56 dot products of length N) into shared memory.
// ---------------------------------------------------------------------------------------------
struct TpsTablesDev {
  const float* Li;
  const float* P_X;
  const float* P_Y;
  const float* grid_X;
  const float* grid_Y;
  int gs;
};

__device__ __forceinline__ void tps_coeffs(const float* __restrict__ theta_b, const TpsTablesDev& t,
                                           float* sQ /*2N*/, float* sW /*2N*/, float* sA /*6*/,
                                           float* sP /*2N*/) {
  const int N = t.gs * t.gs;
  const int L = N + 3;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float px = t.P_X[i], py = t.P_Y[i];
    sP[i] = px;
    sP[N + i] = py;
    sQ[i] = theta_b[i] + px;          // Q_X = theta[:, :N] + P_X_base   (warp.py:207-210)
    sQ[N + i] = theta_b[N + i] + py;  // Q_Y
  }
  __syncthreads();
  for (int r = threadIdx.x; r < 2 * L; r += blockDim.x) {
    const int row = r % L;
    const float* q = sQ + (r / L) * N;
    const float* li = t.Li + row * L;
    float acc = 0.f;
    for (int k = 0; k < N; ++k) acc = fmaf(li[k], q[k], acc);
    if (row < N)
      sW[(r / L) * N + row] = acc;
    else
      sA[(r / L) * 3 + (row - N)] = acc;
  }
  __syncthreads();
}

__device__ __forceinline__ void tps_point(float x, float y, int N, const float* sW, const float* sA,
                                          const float* sP, float& xo, float& yo) {
  float sx = 0.f, sy = 0.f;
#pragma unroll 5
  for (int n = 0; n < N; ++n) {
    float dx = x - sP[n], dy = y - sP[N + n];
    float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));  // torch.pow(.,2)+torch.pow(.,2)
    if (d2 == 0.f) d2 = 1.f;                                     // warp.py:290
    float U = d2 * logf(d2);
    sx = fmaf(sW[n], U, sx);
    sy = fmaf(sW[N + n], U, sy);
  }
  xo = sA[0] + sA[1] * x + sA[2] * y + sx;  // warp.py:303-316
  yo = sA[3] + sA[4] * x + sA[5] * y + sy;
}

// ---------------------------------------------------------------------------------------------
// bilinear sample of one NCHW channel plane; PyTorch grid_sampler semantics, align_corners=False
// ---------------------------------------------------------------------------------------------
struct BilinearTap {
  int x0, y0;            // north-west corner
  float wnw, wne, wsw, wse;
  bool vx0, vx1, vy0, vy1;
};

__device__ __forceinline__ BilinearTap make_tap(float gx, float gy, int Hin, int Win, int padding_mode) {
  float ix = ((gx + 1.f) * Win - 1.f) * 0.5f;
  float iy = ((gy + 1.f) * Hin - 1.f) * 0.5f;
  if (padding_mode == SHINEON_PAD_BORDER) {
    ix = fminf(fmaxf(ix, 0.f), (float)(Win - 1));
    iy = fminf(fmaxf(iy, 0.f), (float)(Hin - 1));
  }
  float fx = floorf(ix), fy = floorf(iy);
  BilinearTap t;
  // weights exactly as ATen's grid_sampler: nw = (ix_se - ix) * (iy_se - iy) ... with ix_se = ix_nw + 1
  float wx1 = ix - fx, wx0 = (fx + 1.f) - ix;
  float wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
  t.wnw = wx0 * wy0;
  t.wne = wx1 * wy0;
  t.wsw = wx0 * wy1;
  t.wse = wx1 * wy1;
  // keep the int conversion defined for wild coordinates (all four taps are out of range there)
  fx = fminf(fmaxf(fx, -2.f), (float)Win + 1.f);
  fy = fminf(fmaxf(fy, -2.f), (float)Hin + 1.f);
  t.x0 = (int)fx;
  t.y0 = (int)fy;
  t.vx0 = t.x0 >= 0 && t.x0 < Win;
  t.vx1 = t.x0 + 1 >= 0 && t.x0 + 1 < Win;
  t.vy0 = t.y0 >= 0 && t.y0 < Hin;
  t.vy1 = t.y0 + 1 >= 0 && t.y0 + 1 < Hin;
  return t;
}

__device__ __forceinline__ float sample_plane(const float* __restrict__ plane, int Win, const BilinearTap& t) {
  float acc = 0.f;
  const float* r0 = plane + (long)t.y0 * Win + t.x0;
  const float* r1 = r0 + Win;
  if (t.vy0 && t.vx0) acc += __ldg(r0) * t.wnw;
  if (t.vy0 && t.vx1) acc += __ldg(r0 + 1) * t.wne;
  if (t.vy1 && t.vx0) acc += __ldg(r1) * t.wsw;
  if (t.vy1 && t.vx1) acc += __ldg(r1 + 1) * t.wse;
  return acc;
}

// Compact tap: 4 clamped offsets + 4 weights, out-of-range taps get weight 0 (x*0 adds exactly 0), so the four
// loads are unconditional for both padding modes and a pixel costs 8 registers.
struct Tap4 {
  int o00, o01, o10, o11;
  float w00, w01, w10, w11;
};
__device__ __forceinline__ Tap4 make_tap4(float gx, float gy, int Hin, int Win, int padding_mode) {
  float ix = ((gx + 1.f) * Win - 1.f) * 0.5f;
  float iy = ((gy + 1.f) * Hin - 1.f) * 0.5f;
  if (padding_mode == SHINEON_PAD_BORDER) {
    ix = fminf(fmaxf(ix, 0.f), (float)(Win - 1));
    iy = fminf(fmaxf(iy, 0.f), (float)(Hin - 1));
  }
  const float fx = floorf(ix), fy = floorf(iy);
  const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix;  // ATen: (ix_se - ix), (ix - ix_nw) ...
  const float wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
  const int x0 = (int)fminf(fmaxf(fx, -2.f), (float)Win + 1.f), y0 = (int)fminf(fmaxf(fy, -2.f), (float)Hin + 1.f);
  const bool vx0 = x0 >= 0 && x0 < Win, vx1 = x0 + 1 >= 0 && x0 + 1 < Win;
  const bool vy0 = y0 >= 0 && y0 < Hin, vy1 = y0 + 1 >= 0 && y0 + 1 < Hin;
  const int cx0 = min(max(x0, 0), Win - 1), cx1 = min(max(x0 + 1, 0), Win - 1);
  const int cy0 = min(max(y0, 0), Hin - 1), cy1 = min(max(y0 + 1, 0), Hin - 1);
  Tap4 t;
  t.o00 = cy0 * Win + cx0; t.o01 = cy0 * Win + cx1; t.o10 = cy1 * Win + cx0; t.o11 = cy1 * Win + cx1;
  t.w00 = (vy0 && vx0) ? wx0 * wy0 : 0.f;
  t.w01 = (vy0 && vx1) ? wx1 * wy0 : 0.f;
  t.w10 = (vy1 && vx0) ? wx0 * wy1 : 0.f;
  t.w11 = (vy1 && vx1) ? wx1 * wy1 : 0.f;
  return t;
}
__device__ __forceinline__ float sample_tap4(const float* __restrict__ plane, const Tap4& t) {
  float acc = __ldg(plane + t.o00) * t.w00;  // nw, ne, sw, se: ATen's accumulation order
  acc += __ldg(plane + t.o01) * t.w01;
  acc += __ldg(plane + t.o10) * t.w10;
  acc += __ldg(plane + t.o11) * t.w11;
  return acc;
}

// Border padding: after clamping the source coordinate every tap index can be clamped too (the weight of an
// out-of-range neighbour is exactly 0), so the four loads are unconditional.
__device__ __forceinline__ float sample_plane_clamped(const float* __restrict__ plane, int Hin, int Win, const BilinearTap& t) {
  const int x0 = min(max(t.x0, 0), Win - 1), x1 = min(max(t.x0 + 1, 0), Win - 1);
  const int y0 = min(max(t.y0, 0), Hin - 1), y1 = min(max(t.y0 + 1, 0), Hin - 1);
  const float* r0 = plane + (long)y0 * Win;
  const float* r1 = plane + (long)y1 * Win;
  float acc = __ldg(r0 + x0) * t.wnw;
  acc += __ldg(r0 + x1) * t.wne;
  acc += __ldg(r1 + x0) * t.wsw;
  acc += __ldg(r1 + x1) * t.wse;
  return acc;
}
__device__ __forceinline__ float sample_any(const float* __restrict__ plane, int Hin, int Win, const BilinearTap& t, int padding_mode) {
  return padding_mode == SHINEON_PAD_BORDER ? sample_plane_clamped(plane, Hin, Win, t) : sample_plane(plane, Win, t);
}

__global__ void __launch_bounds__(256) tps_grid_kernel(const float* __restrict__ theta, TpsTablesDev t,
                                                       float* __restrict__ grid, int H, int W) {
  pdl_grid_sync();
  __shared__ float sQ[2 * kMaxTpsN], sW[2 * kMaxTpsN], sP[2 * kMaxTpsN], sA[6];
  const int b = blockIdx.y;
  const int N = t.gs * t.gs;
  tps_coeffs(theta + (long)b * 2 * N, t, sQ, sW, sA, sP);
  const int HW = H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    int y = p / W, x = p - y * W;
    float xo, yo;
    tps_point(t.grid_X[x], t.grid_Y[y], N, sW, sA, sP, xo, yo);
    reinterpret_cast<float2*>(grid)[(long)b * HW + p] = make_float2(xo, yo);
  }
}

constexpr int kPPT = 4;  // pixels per thread: independent load chains in flight (Little's law, not ALU, bounds these)

__global__ void __launch_bounds__(256)
    grid_sample_kernel(const float* __restrict__ in, const float* __restrict__ grid, float* __restrict__ out,
                       int C, int Hin, int Win, int Hout, int Wout, int padding_mode) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  const int HWo = Hout * Wout;
  const int HWi = Hin * Win;
  const int p0 = blockIdx.x * (256 * kPPT) + threadIdx.x;
  const float2* gb = reinterpret_cast<const float2*>(grid) + (long)b * HWo;
  float2 g[kPPT];
#pragma unroll
  for (int k = 0; k < kPPT; ++k) {
    const int p = p0 + k * 256;
    g[k] = p < HWo ? __ldg(gb + p) : make_float2(0.f, 0.f);
  }
  Tap4 t[kPPT];
#pragma unroll
  for (int k = 0; k < kPPT; ++k) t[k] = make_tap4(g[k].x, g[k].y, Hin, Win, padding_mode);
  for (int c = 0; c < C; ++c) {
    const float* plane = in + ((long)b * C + c) * HWi;
    float v[kPPT];
#pragma unroll
    for (int k = 0; k < kPPT; ++k) v[k] = sample_tap4(plane, t[k]);
#pragma unroll
    for (int k = 0; k < kPPT; ++k) {
      const int p = p0 + k * 256;
      if (p < HWo) out[((long)b * C + c) * HWo + p] = v[k];
    }
  }
}

struct FusedSampleArgs {
  const float* in[3];
  float* out[3];
  int C[3];
  int pad[3];
};

__global__ void __launch_bounds__(256)
    tps_grid_sample_kernel(const float* __restrict__ theta, TpsTablesDev t, FusedSampleArgs a,
                           float* __restrict__ grid_out, int H, int W) {
  pdl_grid_sync();
  __shared__ float sQ[2 * kMaxTpsN], sW[2 * kMaxTpsN], sP[2 * kMaxTpsN], sA[6];
  const int b = blockIdx.y;
  const int N = t.gs * t.gs;
  tps_coeffs(theta + (long)b * 2 * N, t, sQ, sW, sA, sP);
  const int HW = H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    int y = p / W, x = p - y * W;
    float gx, gy;
    tps_point(t.grid_X[x], t.grid_Y[y], N, sW, sA, sP, gx, gy);
    if (grid_out) reinterpret_cast<float2*>(grid_out)[(long)b * HW + p] = make_float2(gx, gy);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (a.in[i] == nullptr) continue;
      BilinearTap tap = make_tap(gx, gy, H, W, a.pad[i]);
      for (int c = 0; c < a.C[i]; ++c)
        a.out[i][((long)b * a.C[i] + c) * HW + p] = sample_plane(a.in[i] + ((long)b * a.C[i] + c) * HW, W, tap);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Batched fused TPS + grid_sample.  U_n(x,y) = d^2 log d^2 depends only on the pixel and the control point, not on
// the image: each thread computes its pixel's N basis values once (N logf) and reuses them for every image of the
// batch chunk, leaving ~2N FMAs + the bilinear gathers per (pixel, image) -> memory-bound instead of logf-bound.
// ---------------------------------------------------------------------------------------------
constexpr int kTpsChunk = 16;  // max images per CTA (fewer for small batches so the grid still fills the GPU)

template <int N>
__global__ void __launch_bounds__(256)
    tps_grid_sample_batched_kernel(const float* __restrict__ theta, TpsTablesDev t, FusedSampleArgs a,
                                   float* __restrict__ grid_out, int B, int H, int W, int chunk) {
  pdl_grid_sync();
  __shared__ float sQ[kTpsChunk][2 * N];
  __shared__ float2 sWxy[kTpsChunk][N];  // (W_X[n], W_Y[n])
  __shared__ float sA[kTpsChunk][6];
  __shared__ float sP[2 * N];
  const int b0 = blockIdx.y * chunk;
  const int nb = min(chunk, B - b0);
  constexpr int L = N + 3;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    sP[i] = t.P_X[i];
    sP[N + i] = t.P_Y[i];
  }
  for (int e = threadIdx.x; e < nb * 2 * N; e += blockDim.x) {
    const int bi = e / (2 * N), k = e - bi * 2 * N;
    sQ[bi][k] = theta[(long)(b0 + bi) * 2 * N + k] + (k < N ? t.P_X[k] : t.P_Y[k - N]);  // warp.py:207-210
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nb * 2 * L; e += blockDim.x) {
    const int bi = e / (2 * L), r = e - bi * 2 * L;
    const int xy = r / L, row = r - xy * L;
    const float* q = sQ[bi] + xy * N;
    const float* li = t.Li + row * L;
    float acc = 0.f;
    for (int k = 0; k < N; ++k) acc = fmaf(li[k], q[k], acc);
    if (row < N) {
      if (xy == 0) sWxy[bi][row].x = acc; else sWxy[bi][row].y = acc;
    } else {
      sA[bi][xy * 3 + (row - N)] = acc;
    }
  }
  __syncthreads();
  const int HW = H * W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const int y = p / W, x = p - y * W;
  const float px = t.grid_X[x], py = t.grid_Y[y];
  float U[N];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const float dx = px - sP[n], dy = py - sP[N + n];
    float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    if (d2 == 0.f) d2 = 1.f;  // warp.py:290
    U[n] = d2 * logf(d2);
  }
  for (int bi = 0; bi < nb; ++bi) {
    float sx = 0.f, sy = 0.f;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const float2 w = sWxy[bi][n];
      sx = fmaf(w.x, U[n], sx);
      sy = fmaf(w.y, U[n], sy);
    }
    const float gx = sA[bi][0] + sA[bi][1] * px + sA[bi][2] * py + sx;  // warp.py:303-316
    const float gy = sA[bi][3] + sA[bi][4] * px + sA[bi][5] * py + sy;
    const int b = b0 + bi;
    if (grid_out) reinterpret_cast<float2*>(grid_out)[(long)b * HW + p] = make_float2(gx, gy);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (a.in[i] == nullptr) continue;
      const Tap4 tap = make_tap4(gx, gy, H, W, a.pad[i]);
      for (int c = 0; c < a.C[i]; ++c)
        a.out[i][((long)b * a.C[i] + c) * HW + p] = sample_tap4(a.in[i] + ((long)b * a.C[i] + c) * HW, tap);
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Fast paths (compile-time channel counts / padding modes).  The generic kernels above are latency- and
// issue-bound (ncu, profiles/r01_memory_ops.md): per-channel loops expose only 16 loads per thread and the runtime
// (C, pad) bookkeeping costs more instructions than the sampling itself.  Here every load of a thread's pixels
// (PPT pixels x C channels x 4 taps) is issued before the first use, offsets are 32-bit, and for border padding
// the neighbour logic collapses to two clamps (the weight of a clamped neighbour is exactly 0).
// ---------------------------------------------------------------------------------------------
struct TapFast {
  // BYTE offsets of the four taps inside one channel plane, unsigned 32-bit: together with a block-uniform 64-bit
  // plane base the loads compile to LDG [R.U32 + UR.64] — no per-tap 64-bit address arithmetic
  unsigned o00, o01, o10, o11;
  float w00, w01, w10, w11;
};
__device__ __forceinline__ void set_offsets(TapFast& t, int o00, int dx, int dy) {
  t.o00 = (unsigned)o00 * 4u;
  t.o01 = t.o00 + (unsigned)dx * 4u;
  t.o10 = t.o00 + (unsigned)dy * 4u;
  t.o11 = t.o10 + (unsigned)dx * 4u;
}
__device__ __forceinline__ float ldg_b(const char* __restrict__ base, unsigned byte_off) {
  return __ldg(reinterpret_cast<const float*>(base + byte_off));
}

template <int PAD>
__device__ __forceinline__ TapFast make_tap_fast(float gx, float gy, int Hin, int Win) {
  float ix = ((gx + 1.f) * Win - 1.f) * 0.5f;
  float iy = ((gy + 1.f) * Hin - 1.f) * 0.5f;
  TapFast t;
  if (PAD == SHINEON_PAD_BORDER) {
    ix = fminf(fmaxf(ix, 0.f), (float)(Win - 1));
    iy = fminf(fmaxf(iy, 0.f), (float)(Hin - 1));
    const float fx = floorf(ix), fy = floorf(iy);
    const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix;
    const float wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
    const int x0 = (int)fx, y0 = (int)fy;  // in [0, W-1] x [0, H-1]
    // at the border wx1 == 0 exactly: any finite neighbour value contributes 0
    set_offsets(t, y0 * Win + x0, x0 + 1 < Win ? 1 : 0, y0 + 1 < Hin ? Win : 0);
    t.w00 = wx0 * wy0; t.w01 = wx1 * wy0; t.w10 = wx0 * wy1; t.w11 = wx1 * wy1;
  } else {
    const float fx = floorf(ix), fy = floorf(iy);
    const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix;
    const float wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
    const int x0 = (int)fminf(fmaxf(fx, -2.f), (float)Win + 1.f), y0 = (int)fminf(fmaxf(fy, -2.f), (float)Hin + 1.f);
    const bool vx0 = (unsigned)x0 < (unsigned)Win, vx1 = (unsigned)(x0 + 1) < (unsigned)Win;
    const bool vy0 = (unsigned)y0 < (unsigned)Hin, vy1 = (unsigned)(y0 + 1) < (unsigned)Hin;
    const int cx0 = min(max(x0, 0), Win - 1), cy0 = min(max(y0, 0), Hin - 1);
    set_offsets(t, cy0 * Win + cx0, (vx0 && vx1) ? 1 : 0, (vy0 && vy1) ? Win : 0);
    // a tap that is out of range gets weight 0; its (clamped) address still points inside the plane
    t.w00 = (vy0 && vx0) ? wx0 * wy0 : 0.f;
    t.w01 = (vy0 && vx1) ? wx1 * wy0 : 0.f;
    t.w10 = (vy1 && vx0) ? wx0 * wy1 : 0.f;
    t.w11 = (vy1 && vx1) ? wx1 * wy1 : 0.f;
    // when exactly one of a neighbour pair is valid the valid one must sit at the clamped address: x0 == -1 -> cx0 = 0
    // is the "x1" tap (weight w01) and dx = 0 makes both loads hit it; same for y0 == -1
  }
  return t;
}

// Loads the 4 taps of C channel planes (stride `plane_stride` elements) for one pixel.
template <int C>
__device__ __forceinline__ void load_taps(const float* __restrict__ base, long plane_stride, const TapFast& t, float (&v)[C][4]) {
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const char* pl = reinterpret_cast<const char*>(base + c * plane_stride);  // block-uniform
    v[c][0] = ldg_b(pl, t.o00);
    v[c][1] = ldg_b(pl, t.o01);
    v[c][2] = ldg_b(pl, t.o10);
    v[c][3] = ldg_b(pl, t.o11);
  }
}
__device__ __forceinline__ void st_b(float* __restrict__ base, unsigned byte_off, float v) {
  *reinterpret_cast<float*>(reinterpret_cast<char*>(base) + byte_off) = v;
}
__device__ __forceinline__ float blend_taps(const float (&v)[4], const TapFast& t) {
  float acc = v[0] * t.w00;  // nw, ne, sw, se: ATen's accumulation order
  acc += v[1] * t.w01;
  acc += v[2] * t.w10;
  acc += v[3] * t.w11;
  return acc;
}

template <int C, int PAD, int PPT>
__global__ void __launch_bounds__(256)
    grid_sample_fast_kernel(const float* __restrict__ in, const float* __restrict__ grid, float* __restrict__ out,
                            int Hin, int Win, int HWo) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  const int HWi = Hin * Win;
  const int p0 = blockIdx.x * (256 * PPT) + threadIdx.x;
  const float2* gb = reinterpret_cast<const float2*>(grid) + (long)b * HWo;
  const float* ib = in + (long)b * C * HWi;
  float* ob = out + (long)b * C * HWo;
  float2 g[PPT];
#pragma unroll
  for (int k = 0; k < PPT; ++k) g[k] = __ldg(gb + min(p0 + k * 256, HWo - 1));
  TapFast t[PPT];
  float v[PPT][C][4];
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    t[k] = make_tap_fast<PAD>(g[k].x, g[k].y, Hin, Win);
    load_taps<C>(ib, HWi, t[k], v[k]);
  }
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int p = p0 + k * 256;
    if (p < HWo) {
#pragma unroll
      for (int c = 0; c < C; ++c) st_b(ob + (long)c * HWo, (unsigned)p * 4u, blend_taps(v[k][c], t[k]));
    }
  }
}

template <int C, int PPT>
__global__ void __launch_bounds__(256)
    resample2d_fast_kernel(const float* __restrict__ in1, const float* __restrict__ flow, float* __restrict__ out,
                           int H, int W) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  const int HW = H * W;
  const int p0 = blockIdx.x * (256 * PPT) + threadIdx.x;
  const float* fb = flow + (long)b * 2 * HW;
  const float* ib = in1 + (long)b * C * HW;
  float* ob = out + (long)b * C * HW;
  float dx[PPT], dy[PPT];
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int p = min(p0 + k * 256, HW - 1);
    dx[k] = __ldg(fb + p);
    dy[k] = __ldg(fb + HW + p);
  }
  TapFast t[PPT];
  float v[PPT][C][4];
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int p = min(p0 + k * 256, HW - 1);
    const int y = p / W, x = p - y * W;
    const float xf = (float)x + dx[k], yf = (float)y + dy[k];
    const float fx = floorf(xf), fy = floorf(yf);
    const float alpha = xf - fx, beta = yf - fy;  // resample2d_kernel.cu:42-43
    const float cfx = fminf(fmaxf(fx, -4.f), (float)W + 4.f), cfy = fminf(fmaxf(fy, -4.f), (float)H + 4.f);
    const int xL = max(min((int)cfx, W - 1), 0), xR = max(min((int)cfx + 1, W - 1), 0);
    const int yT = max(min((int)cfy, H - 1), 0), yB = max(min((int)cfy + 1, H - 1), 0);
    set_offsets(t[k], yT * W + xL, xR - xL, (yB - yT) * W);
    t[k].w00 = (1.f - alpha) * (1.f - beta); t[k].w01 = alpha * (1.f - beta);
    t[k].w10 = (1.f - alpha) * beta; t[k].w11 = alpha * beta;
    load_taps<C>(ib, HW, t[k], v[k]);
  }
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int p = p0 + k * 256;
    if (p < HW) {
#pragma unroll
      for (int c = 0; c < C; ++c) {  // same accumulation order as resample2d_kernel.cu:56-59 (0 + w*v ...)
        float a = 0.f;
        a += t[k].w00 * v[k][c][0];
        a += t[k].w01 * v[k][c][1];
        a += t[k].w10 * v[k][c][2];
        a += t[k].w11 * v[k][c][3];
        st_b(ob + (long)c * HW, (unsigned)p * 4u, a);
      }
    }
  }
}

// Batched fused TPS + grid_sample, specialised: N control points, up to three inputs with compile-time channel
// counts / padding (C == 0: absent).  Two pixels per thread share every (W_X, W_Y) shared-memory read; the two
// coordinate sums advance with one packed FFMA2 per control point.
template <int N, int C0, int P0, int C1, int P1, int C2, int P2>
__global__ void __launch_bounds__(128)
    tps_grid_sample_fast_kernel(const float* __restrict__ theta, TpsTablesDev t, FusedSampleArgs a,
                                float* __restrict__ grid_out, int B, int H, int W, int chunk) {
  pdl_grid_sync();
  __shared__ float sQ[kTpsChunk][2 * N];
  __shared__ float2 sWxy[kTpsChunk][N];
  __shared__ float sA[kTpsChunk][6];
  __shared__ float sP[2 * N];
  const int b0 = blockIdx.y * chunk;
  const int nb = min(chunk, B - b0);
  constexpr int L = N + 3;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    sP[i] = t.P_X[i];
    sP[N + i] = t.P_Y[i];
  }
  for (int e = threadIdx.x; e < nb * 2 * N; e += blockDim.x) {
    const int bi = e / (2 * N), k = e - bi * 2 * N;
    sQ[bi][k] = theta[(long)(b0 + bi) * 2 * N + k] + (k < N ? t.P_X[k] : t.P_Y[k - N]);  // warp.py:207-210
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nb * 2 * L; e += blockDim.x) {
    const int bi = e / (2 * L), r = e - bi * 2 * L;
    const int xy = r / L, row = r - xy * L;
    const float* q = sQ[bi] + xy * N;
    const float* li = t.Li + row * L;
    float acc = 0.f;
    for (int k = 0; k < N; ++k) acc = fmaf(li[k], q[k], acc);
    if (row < N) {
      if (xy == 0) sWxy[bi][row].x = acc; else sWxy[bi][row].y = acc;
    } else {
      sA[bi][xy * 3 + (row - N)] = acc;
    }
  }
  __syncthreads();
  const int HW = H * W;
  // two pixels per thread, 128 apart (coalesced per warp)
  const int pA = blockIdx.x * 256 + threadIdx.x, pB = pA + 128;
  if (pA >= HW) return;
  const bool hasB = pB < HW;
  const int pBc = hasB ? pB : pA;
  const int yA = pA / W, xA = pA - yA * W, yB = pBc / W, xB = pBc - yB * W;
  const float pxA = t.grid_X[xA], pyA = t.grid_Y[yA], pxB = t.grid_X[xB], pyB = t.grid_Y[yB];
  float UA[N], UB[N];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const float cx = sP[n], cy = sP[N + n];
    float dx = pxA - cx, dy = pyA - cy;
    float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    if (d2 == 0.f) d2 = 1.f;  // warp.py:290
    UA[n] = d2 * logf(d2);
    dx = pxB - cx; dy = pyB - cy;
    d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    if (d2 == 0.f) d2 = 1.f;
    UB[n] = d2 * logf(d2);
  }
  for (int bi = 0; bi < nb; ++bi) {
    float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const float2 w = sWxy[bi][n];
      sa = __ffma2_rn(w, make_float2(UA[n], UA[n]), sa);  // (sx, sy) += (W_X[n], W_Y[n]) * U_n: same fp32 FMAs, one issue
      sb = __ffma2_rn(w, make_float2(UB[n], UB[n]), sb);
    }
    const float a0 = sA[bi][0], a1 = sA[bi][1], a2 = sA[bi][2], a3 = sA[bi][3], a4 = sA[bi][4], a5 = sA[bi][5];
    const float gxA = a0 + a1 * pxA + a2 * pyA + sa.x, gyA = a3 + a4 * pxA + a5 * pyA + sa.y;  // warp.py:303-316
    const float gxB = a0 + a1 * pxB + a2 * pyB + sb.x, gyB = a3 + a4 * pxB + a5 * pyB + sb.y;
    const int b = b0 + bi;
    if (grid_out) {
      reinterpret_cast<float2*>(grid_out)[(long)b * HW + pA] = make_float2(gxA, gyA);
      if (hasB) reinterpret_cast<float2*>(grid_out)[(long)b * HW + pB] = make_float2(gxB, gyB);
    }
    // all loads of both pixels first, then blends + stores
    float v0[2][C0 > 0 ? C0 : 1][4], v1[2][C1 > 0 ? C1 : 1][4], v2[2][C2 > 0 ? C2 : 1][4];
    TapFast t0[2], t1[2], t2[2];
    if (C0 > 0) {
      t0[0] = make_tap_fast<P0>(gxA, gyA, H, W); t0[1] = make_tap_fast<P0>(gxB, gyB, H, W);
      load_taps<(C0 > 0 ? C0 : 1)>(a.in[0] + (long)b * C0 * HW, HW, t0[0], v0[0]);
      load_taps<(C0 > 0 ? C0 : 1)>(a.in[0] + (long)b * C0 * HW, HW, t0[1], v0[1]);
    }
    if (C1 > 0) {
      if (P1 == P0 && C0 > 0) { t1[0] = t0[0]; t1[1] = t0[1]; }
      else { t1[0] = make_tap_fast<P1>(gxA, gyA, H, W); t1[1] = make_tap_fast<P1>(gxB, gyB, H, W); }
      load_taps<(C1 > 0 ? C1 : 1)>(a.in[1] + (long)b * C1 * HW, HW, t1[0], v1[0]);
      load_taps<(C1 > 0 ? C1 : 1)>(a.in[1] + (long)b * C1 * HW, HW, t1[1], v1[1]);
    }
    if (C2 > 0) {
      if (P2 == P1 && C1 > 0) { t2[0] = t1[0]; t2[1] = t1[1]; }
      else if (P2 == P0 && C0 > 0) { t2[0] = t0[0]; t2[1] = t0[1]; }
      else { t2[0] = make_tap_fast<P2>(gxA, gyA, H, W); t2[1] = make_tap_fast<P2>(gxB, gyB, H, W); }
      load_taps<(C2 > 0 ? C2 : 1)>(a.in[2] + (long)b * C2 * HW, HW, t2[0], v2[0]);
      load_taps<(C2 > 0 ? C2 : 1)>(a.in[2] + (long)b * C2 * HW, HW, t2[1], v2[1]);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (k == 1 && !hasB) break;
      const unsigned pb = (unsigned)(k == 0 ? pA : pB) * 4u;
      if (C0 > 0) {
#pragma unroll
        for (int c = 0; c < C0; ++c) st_b(a.out[0] + ((long)b * C0 + c) * HW, pb, blend_taps(v0[k][c], t0[k]));
      }
      if (C1 > 0) {
#pragma unroll
        for (int c = 0; c < C1; ++c) st_b(a.out[1] + ((long)b * C1 + c) * HW, pb, blend_taps(v1[k][c], t1[k]));
      }
      if (C2 > 0) {
#pragma unroll
        for (int c = 0; c < C2; ++c) st_b(a.out[2] + ((long)b * C2 + c) * HW, pb, blend_taps(v2[k][c], t2[k]));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TPS warp of the cloth for the uint8 pipeline (TryOnPipeline.run_raw): the same batched TPS + grid_sample(border) as
// tps_grid_sample_fast_kernel<N, 3, BORDER>, but
//   * the source is the decoded 8-bit cloth [B,H,W,3] (channel-last), normalised on load exactly like
//     ToTensor + Normalize(0.5, 0.5) do (frame_prep.cu norm_u8), so the f32 cloth tensor is never materialised;
//   * besides the f32 NCHW warped cloth (read again by the compose kernel) every sample is also written, split into 16-bit
//     hi/lo halves, into the cloth channels of the U-Net stem's space-to-depth planes
//       z[b][Y][X][(py*2+px)*ctot + c_off + c] = x[c][2Y-1+py][2X-1+px]   (norm_act.cu nchw_s2d_planes_kernel)
//     whose person channels the frame-prep kernel filled: torch.cat + the layout pass disappear from the step.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float norm_u8_tps(uint8_t u) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.f), 0.5f), 0.5f);
}

template <int N>
__global__ void __launch_bounds__(128)
    tps_warp_u8_planes_kernel(const float* __restrict__ theta, TpsTablesDev t, const uint8_t* __restrict__ cloth,
                              float* __restrict__ out, plane_t* __restrict__ zh, plane_t* __restrict__ zl, int zc, int ctot,
                              int c_off, int fmt, int B, int H, int W, int chunk) {
  pdl_grid_sync();
  __shared__ float sQ[kTpsChunk][2 * N];
  __shared__ float2 sWxy[kTpsChunk][N];
  __shared__ float sA[kTpsChunk][6];
  __shared__ float sP[2 * N];
  __shared__ float s_lut[256];  // ToTensor + Normalize of every byte value (two IEEE divisions each: once per CTA, not per tap)
  const int b0 = blockIdx.y * chunk;
  const int nb = min(chunk, B - b0);
  constexpr int L = N + 3;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = norm_u8_tps((uint8_t)i);
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    sP[i] = t.P_X[i];
    sP[N + i] = t.P_Y[i];
  }
  for (int e = threadIdx.x; e < nb * 2 * N; e += blockDim.x) {
    const int bi = e / (2 * N), k = e - bi * 2 * N;
    sQ[bi][k] = theta[(long)(b0 + bi) * 2 * N + k] + (k < N ? t.P_X[k] : t.P_Y[k - N]);  // warp.py:207-210
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nb * 2 * L; e += blockDim.x) {
    const int bi = e / (2 * L), r = e - bi * 2 * L;
    const int xy = r / L, row = r - xy * L;
    const float* q = sQ[bi] + xy * N;
    const float* li = t.Li + row * L;
    float acc = 0.f;
    for (int k = 0; k < N; ++k) acc = fmaf(li[k], q[k], acc);
    if (row < N) {
      if (xy == 0) sWxy[bi][row].x = acc; else sWxy[bi][row].y = acc;
    } else {
      sA[bi][xy * 3 + (row - N)] = acc;
    }
  }
  __syncthreads();
  const int HW = H * W;
  const int pA = blockIdx.x * 256 + threadIdx.x, pB = pA + 128;
  if (pA >= HW) return;
  const bool hasB = pB < HW;
  const int pBc = hasB ? pB : pA;
  const int yy[2] = {pA / W, pBc / W};
  const int xx[2] = {pA - yy[0] * W, pBc - yy[1] * W};
  const float pxA = t.grid_X[xx[0]], pyA = t.grid_Y[yy[0]], pxB = t.grid_X[xx[1]], pyB = t.grid_Y[yy[1]];
  float UA[N], UB[N];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const float cx = sP[n], cy = sP[N + n];
    float dx = pxA - cx, dy = pyA - cy;
    float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    if (d2 == 0.f) d2 = 1.f;  // warp.py:290
    UA[n] = d2 * logf(d2);
    dx = pxB - cx; dy = pyB - cy;
    d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    if (d2 == 0.f) d2 = 1.f;
    UB[n] = d2 * logf(d2);
  }
  // where the two pixels land in the space-to-depth planes (element offsets of channel 0 of the cloth group)
  const int Wz = W / 2 + 1, Hz = H / 2 + 1;
  long zoff[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int Y = (yy[k] + 1) >> 1, py = (yy[k] + 1) & 1, X = (xx[k] + 1) >> 1, px = (xx[k] + 1) & 1;
    zoff[k] = ((long)Y * Wz + X) * zc + (py * 2 + px) * ctot + c_off;
  }
  for (int bi = 0; bi < nb; ++bi) {
    float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const float2 w = sWxy[bi][n];
      sa = __ffma2_rn(w, make_float2(UA[n], UA[n]), sa);
      sb = __ffma2_rn(w, make_float2(UB[n], UB[n]), sb);
    }
    const float a0 = sA[bi][0], a1 = sA[bi][1], a2 = sA[bi][2], a3 = sA[bi][3], a4 = sA[bi][4], a5 = sA[bi][5];
    const float gx[2] = {a0 + a1 * pxA + a2 * pyA + sa.x, a0 + a1 * pxB + a2 * pyB + sb.x};  // warp.py:303-316
    const float gy[2] = {a3 + a4 * pxA + a5 * pyA + sa.y, a3 + a4 * pxB + a5 * pyB + sb.y};
    const int b = b0 + bi;
    const uint8_t* src = cloth + (long)b * HW * 3;
    TapFast tp[2];
    uint8_t u[2][4][3];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      tp[k] = make_tap_fast<SHINEON_PAD_BORDER>(gx[k], gy[k], H, W);  // byte offsets of f32 planes: /4 = pixel index
      const unsigned o[4] = {tp[k].o00 >> 2, tp[k].o01 >> 2, tp[k].o10 >> 2, tp[k].o11 >> 2};
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c) u[k][j][c] = __ldg(src + o[j] * 3u + c);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (k == 1 && !hasB) break;
      const int p = k == 0 ? pA : pB;
      plane_t* zhb = zh + (long)b * Hz * Wz * zc + zoff[k];
      plane_t* zlb = zl ? zl + (long)b * Hz * Wz * zc + zoff[k] : nullptr;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v[4] = {s_lut[u[k][0][c]], s_lut[u[k][1][c]], s_lut[u[k][2][c]], s_lut[u[k][3][c]]};
        const float r = blend_taps(v, tp[k]);
        out[((long)b * 3 + c) * HW + p] = r;
        plane_t hh, ll;
        split16(r, fmt, hh, ll);
        zhb[c] = hh;
        if (zlb) zlb[c] = ll;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Resample2d  (resample2d_kernel.cu).  kernel_size == 1 (the only value the reference uses).
// One thread per output pixel; the flow is read once and reused for every channel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    resample2d_fwd_kernel(const float* __restrict__ in1, const float* __restrict__ flow,
                          float* __restrict__ out, int C, int H, int W, int bilinear) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  const int HW = H * W;
  const int p0 = blockIdx.x * (256 * kPPT) + threadIdx.x;
  float dx[kPPT], dy[kPPT];
#pragma unroll
  for (int k = 0; k < kPPT; ++k) {
    const int p = min(p0 + k * 256, HW - 1);
    dx[k] = __ldg(flow + ((long)b * 2 + 0) * HW + p);
    dy[k] = __ldg(flow + ((long)b * 2 + 1) * HW + p);
  }
  int o00[kPPT], o01[kPPT], o10[kPPT], o11[kPPT];
  float w00[kPPT], w01[kPPT], w10[kPPT], w11[kPPT];
#pragma unroll
  for (int k = 0; k < kPPT; ++k) {
    const int p = min(p0 + k * 256, HW - 1);
    const int y = p / W, x = p - y * W;
    const float xf = (float)x + dx[k], yf = (float)y + dy[k];
    if (bilinear) {
      const float fx = floorf(xf), fy = floorf(yf);
      const float alpha = xf - fx, beta = yf - fy;  // resample2d_kernel.cu:42-43
      // int(floor(xf)) with a defined result for huge flows
      const float cfx = fminf(fmaxf(fx, -4.f), (float)W + 4.f), cfy = fminf(fmaxf(fy, -4.f), (float)H + 4.f);
      const int xL = max(min((int)cfx, W - 1), 0), xR = max(min((int)cfx + 1, W - 1), 0);
      const int yT = max(min((int)cfy, H - 1), 0), yB = max(min((int)cfy + 1, H - 1), 0);
      o00[k] = yT * W + xL; o01[k] = yT * W + xR; o10[k] = yB * W + xL; o11[k] = yB * W + xR;
      w00[k] = (1.f - alpha) * (1.f - beta); w01[k] = alpha * (1.f - beta);
      w10[k] = (1.f - alpha) * beta; w11[k] = alpha * beta;
    } else {
      const float nx = fminf(fmaxf(floorf(xf + 0.5f), -4.f), (float)W + 4.f);
      const float ny = fminf(fmaxf(floorf(yf + 0.5f), -4.f), (float)H + 4.f);
      o00[k] = max(min((int)ny, H - 1), 0) * W + max(min((int)nx, W - 1), 0);
      o01[k] = o10[k] = o11[k] = o00[k];
      w00[k] = 1.f; w01[k] = w10[k] = w11[k] = 0.f;
    }
  }
  for (int c = 0; c < C; ++c) {
    const float* pl = in1 + ((long)b * C + c) * HW;
    float v[kPPT];
#pragma unroll
    for (int k = 0; k < kPPT; ++k) {
      if (bilinear) {  // same accumulation order as resample2d_kernel.cu:56-59
        float a = 0.f;
        a += w00[k] * __ldg(pl + o00[k]);
        a += w01[k] * __ldg(pl + o01[k]);
        a += w10[k] * __ldg(pl + o10[k]);
        a += w11[k] * __ldg(pl + o11[k]);
        v[k] = a;
      } else {
        v[k] = __ldg(pl + o00[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < kPPT; ++k) {
      const int p = p0 + k * 256;
      if (p < HW) out[((long)b * C + c) * HW + p] = v[k];
    }
  }
}

// d/d in1: atomic scatter (resample2d_kernel.cu:76-125).  NB the reference takes alpha/beta from
// `xf - int(xf)` (truncation, :105-106) here, unlike the forward's floor; reproduced on purpose.
__global__ void __launch_bounds__(256)
    resample2d_bwd_in1_kernel(const float* __restrict__ flow, const float* __restrict__ gout,
                              float* __restrict__ gin1, int C, int H, int W) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  const int HW = H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    int y = p / W, x = p - y * W;
    float dx = flow[((long)b * 2 + 0) * HW + p];
    float dy = flow[((long)b * 2 + 1) * HW + p];
    float xf = (float)x + dx, yf = (float)y + dy;
    float cxf = fminf(fmaxf(xf, -1e9f), 1e9f), cyf = fminf(fmaxf(yf, -1e9f), 1e9f);
    float alpha = xf - (float)(int)cxf, beta = yf - (float)(int)cyf;
    float cfx = fminf(fmaxf(floorf(xf), -4.f), (float)W + 4.f), cfy = fminf(fmaxf(floorf(yf), -4.f), (float)H + 4.f);
    int xL = max(min((int)cfx, W - 1), 0), xR = max(min((int)cfx + 1, W - 1), 0);
    int yT = max(min((int)cfy, H - 1), 0), yB = max(min((int)cfy + 1, H - 1), 0);
    for (int c = 0; c < C; ++c) {
      float g = gout[((long)b * C + c) * HW + p];
      float* pl = gin1 + ((long)b * C + c) * HW;
      atomicAdd(pl + yT * W + xL, (1.f - alpha) * (1.f - beta) * g);
      atomicAdd(pl + yT * W + xR, alpha * (1.f - beta) * g);
      atomicAdd(pl + yB * W + xL, (1.f - alpha) * beta * g);
      atomicAdd(pl + yB * W + xR, alpha * beta * g);
    }
  }
}

// d/d flow (resample2d_kernel.cu:128-198): channel 0 (even) = d/dx, channel 1 (odd) = d/dy.
__global__ void __launch_bounds__(256)
    resample2d_bwd_flow_kernel(const float* __restrict__ in1, const float* __restrict__ flow,
                               const float* __restrict__ gout, float* __restrict__ gflow, int C, int H, int W) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  const int HW = H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    int y = p / W, x = p - y * W;
    float dx = flow[((long)b * 2 + 0) * HW + p];
    float dy = flow[((long)b * 2 + 1) * HW + p];
    float xf = (float)x + dx, yf = (float)y + dy;
    float fx = floorf(xf), fy = floorf(yf);
    float cfx = fminf(fmaxf(fx, -4.f), (float)W + 4.f), cfy = fminf(fmaxf(fy, -4.f), (float)H + 4.f);
    int xL = max(min((int)cfx, W - 1), 0), xR = max(min((int)cfx + 1, W - 1), 0);
    int yT = max(min((int)cfy, H - 1), 0), yB = max(min((int)cfy + 1, H - 1), 0);
    float gx_gamma = 1.f - (yf - fy);  // even channel (:181-192): weights along y, difference along x
    float gy_gamma = 1.f - (xf - fx);  // odd channel  (:168-179): weights along x, difference along y
    float ox = 0.f, oy = 0.f;
    for (int c = 0; c < C; ++c) {
      const float* pl = in1 + ((long)b * C + c) * HW;
      float g = gout[((long)b * C + c) * HW + p];
      float vTL = __ldg(pl + yT * W + xL), vTR = __ldg(pl + yT * W + xR);
      float vBL = __ldg(pl + yB * W + xL), vBR = __ldg(pl + yB * W + xR);
      ox += gx_gamma * g * vTR;
      ox -= gx_gamma * g * vTL;
      ox += (1.f - gx_gamma) * g * vBR;
      ox -= (1.f - gx_gamma) * g * vBL;
      oy += gy_gamma * g * vBL;
      oy -= gy_gamma * g * vTL;
      oy += (1.f - gy_gamma) * g * vBR;
      oy -= (1.f - gy_gamma) * g * vTR;
    }
    gflow[((long)b * 2 + 0) * HW + p] = ox;
    gflow[((long)b * 2 + 1) * HW + p] = oy;
  }
}

// ---------------------------------------------------------------------------------------------
// ChannelNorm (channelnorm_kernel.cu): out = sqrt(sum_c x^2); norm_deg is ignored by the reference.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    channelnorm_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int HW) {
  pdl_grid_sync();
  // grid.y = image; each thread owns 4 consecutive pixels (float4 per channel plane) when HW % 4 == 0
  const int b = blockIdx.y;
  const float* ib = in + (long)b * C * HW;
  float* ob = out + (long)b * HW;
  if ((HW & 3) == 0) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q * 4 >= HW) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < C; ++c) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(ib + (long)c * HW) + q);
      acc.x = fmaf(v.x, v.x, acc.x); acc.y = fmaf(v.y, v.y, acc.y);
      acc.z = fmaf(v.z, v.z, acc.z); acc.w = fmaf(v.w, v.w, acc.w);
    }
    reinterpret_cast<float4*>(ob)[q] = make_float4(sqrtf(acc.x), sqrtf(acc.y), sqrtf(acc.z), sqrtf(acc.w));
  } else {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
      float acc = 0.f;
      for (int c = 0; c < C; ++c) {
        const float v = ib[(long)c * HW + p];
        acc = fmaf(v, v, acc);
      }
      ob[p] = sqrtf(acc);
    }
  }
}

__global__ void __launch_bounds__(256)
    channelnorm_bwd_kernel(const float* __restrict__ in, const float* __restrict__ out,
                           const float* __restrict__ gout, float* __restrict__ gin, int C, long HW, long total) {
  pdl_grid_sync();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long b = i / (C * HW), p = i % HW;
    long o = b * HW + p;
    // channelnorm_kernel.cu:93: float*float / (float + 1e-9 (double))
    double den = (double)out[o] + 1e-9;
    gin[i] = (float)((double)(gout[o] * in[i]) / den);
  }
}

static inline TpsTablesDev to_dev(const shineon_tps_tables* t) {
  TpsTablesDev d{t->Li, t->P_X, t->P_Y, t->grid_X, t->grid_Y, t->grid_size};
  return d;
}

static inline dim3 pixel_grid(int HW, int B) {
  int bx = cdiv(HW, 256);
  if (bx > 4096) bx = 4096;
  return dim3(bx, B);
}

}  // namespace shineon

using namespace shineon;

extern "C" int shineon_tps_grid_fwd(const float* theta, const shineon_tps_tables* tps, float* grid, int B, int H,
                                    int W, shineon_stream_t stream) {
  SHINEON_REQUIRE(theta && tps && grid, "tps_grid: null pointer");
  SHINEON_REQUIRE(tps->Li && tps->P_X && tps->P_Y && tps->grid_X && tps->grid_Y, "tps_grid: null table");
  SHINEON_REQUIRE(tps->grid_size >= 2 && tps->grid_size * tps->grid_size <= kMaxTpsN, "tps_grid: grid_size %d unsupported", tps->grid_size);
  SHINEON_REQUIRE(B >= 0 && H > 0 && W > 0 && B <= 65535, "tps_grid: bad shape");
  if (B == 0) return SHINEON_OK;
  const int N = tps->grid_size * tps->grid_size;
  if (N == 25 || N == 9) {
    FusedSampleArgs a;
    for (int i = 0; i < 3; ++i) { a.in[i] = nullptr; a.out[i] = nullptr; a.C[i] = 0; a.pad[i] = 0; }
    const int chunk = B >= 128 ? kTpsChunk : (B >= 32 ? 8 : 2);
    dim3 g(cdiv(H * W, 256), cdiv(B, chunk));
    if (N == 25)
      klaunch(tps_grid_sample_batched_kernel<25>, g, 256, 0, (cudaStream_t)stream, theta, to_dev(tps), a, grid, B, H, W, chunk);
    else
      klaunch(tps_grid_sample_batched_kernel<9>, g, 256, 0, (cudaStream_t)stream, theta, to_dev(tps), a, grid, B, H, W, chunk);
    return after_launch("tps_grid_sample_batched_kernel");
  }
  klaunch(tps_grid_kernel, pixel_grid(H * W, B), 256, 0, (cudaStream_t)stream, theta, to_dev(tps), grid, H, W);
  return after_launch("tps_grid_kernel");
}

extern "C" int shineon_grid_sample_fwd(const float* input, const float* grid, float* out, int B, int C, int Hin,
                                       int Win, int Hout, int Wout, int padding_mode, shineon_stream_t stream) {
  SHINEON_REQUIRE(input && grid && out, "grid_sample: null pointer");
  SHINEON_REQUIRE(padding_mode == SHINEON_PAD_ZEROS || padding_mode == SHINEON_PAD_BORDER, "grid_sample: padding_mode %d", padding_mode);
  SHINEON_REQUIRE(B >= 0 && C > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0 && B <= 65535, "grid_sample: bad shape");
  if (B == 0) return SHINEON_OK;
  if (C <= 4 && (long)C * Hin * Win < (1l << 31)) {
    constexpr int PPT = 2;  // measured on B200 (B=512): PPT 1 / 2 / 4 -> 0.64 / 0.68 / 0.53 of HBM peak
    const dim3 g(cdiv(Hout * Wout, 256 * PPT), B);
    cudaStream_t st = (cudaStream_t)stream;
#define SHINEON_GS(C_)                                                                                                       \
  if (padding_mode == SHINEON_PAD_BORDER)                                                                                   \
    klaunch(grid_sample_fast_kernel<C_, SHINEON_PAD_BORDER, PPT>, g, 256, 0, st, input, grid, out, Hin, Win, Hout * Wout);       \
  else                                                                                                                      \
    klaunch(grid_sample_fast_kernel<C_, SHINEON_PAD_ZEROS, PPT>, g, 256, 0, st, input, grid, out, Hin, Win, Hout * Wout)
    switch (C) {
      case 1: SHINEON_GS(1); break;
      case 2: SHINEON_GS(2); break;
      case 3: SHINEON_GS(3); break;
      default: SHINEON_GS(4); break;
    }
#undef SHINEON_GS
    return after_launch("grid_sample_fast_kernel");
  }
  klaunch(grid_sample_kernel, dim3(cdiv(Hout * Wout, 256 * kPPT), B), 256, 0, (cudaStream_t)stream, input, grid, out, C, Hin, Win,
                                                                                              Hout, Wout, padding_mode);
  return after_launch("grid_sample_kernel");
}

extern "C" int shineon_tps_grid_sample_fwd(const float* theta, const shineon_tps_tables* tps, int B, int H, int W,
                                           const float* in0, int C0, int pad0, float* out0, const float* in1,
                                           int C1, int pad1, float* out1, const float* in2, int C2, int pad2,
                                           float* out2, float* grid_out, shineon_stream_t stream) {
  SHINEON_REQUIRE(theta && tps, "tps_grid_sample: null pointer");
  SHINEON_REQUIRE(tps->Li && tps->P_X && tps->P_Y && tps->grid_X && tps->grid_Y, "tps_grid_sample: null table");
  SHINEON_REQUIRE(tps->grid_size >= 2 && tps->grid_size * tps->grid_size <= kMaxTpsN, "tps_grid_sample: grid_size %d unsupported", tps->grid_size);
  SHINEON_REQUIRE(B >= 0 && H > 0 && W > 0 && B <= 65535, "tps_grid_sample: bad shape");
  SHINEON_REQUIRE((in0 == nullptr) == (out0 == nullptr) && (in1 == nullptr) == (out1 == nullptr) && (in2 == nullptr) == (out2 == nullptr), "tps_grid_sample: in/out mismatch");
  if (B == 0) return SHINEON_OK;
  FusedSampleArgs a;
  a.in[0] = in0; a.out[0] = out0; a.C[0] = C0; a.pad[0] = pad0;
  a.in[1] = in1; a.out[1] = out1; a.C[1] = C1; a.pad[1] = pad1;
  a.in[2] = in2; a.out[2] = out2; a.C[2] = C2; a.pad[2] = pad2;
  const int N = tps->grid_size * tps->grid_size;
  if (N == 25) {  // the ShineOn configuration (grid_size 5) with the two input signatures WarpModel uses
    const int chunk = B >= 128 ? kTpsChunk : (B >= 32 ? 8 : 2);
    dim3 grid(cdiv(H * W, 256), cdiv(B, chunk));
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int BD = SHINEON_PAD_BORDER, ZR = SHINEON_PAD_ZEROS;
    const bool i0 = in0 && C0 == 3 && pad0 == BD;
    if (i0 && !in1 && !in2) {
      klaunch(tps_grid_sample_fast_kernel<25, 3, BD, 0, 0, 0, 0>, grid, 128, 0, st, theta, to_dev(tps), a, grid_out, B, H, W, chunk);
      return after_launch("tps_grid_sample_fast_kernel");
    }
    if (i0 && in1 && C1 == 3 && pad1 == ZR && !in2) {  // cloth + an RGB "mask" / grid image
      klaunch(tps_grid_sample_fast_kernel<25, 3, BD, 3, ZR, 0, 0>, grid, 128, 0, st, theta, to_dev(tps), a, grid_out, B, H, W, chunk);
      return after_launch("tps_grid_sample_fast_kernel");
    }
    if (i0 && in1 && C1 == 1 && pad1 == ZR && !in2) {
      klaunch(tps_grid_sample_fast_kernel<25, 3, BD, 1, ZR, 0, 0>, grid, 128, 0, st, theta, to_dev(tps), a, grid_out, B, H, W, chunk);
      return after_launch("tps_grid_sample_fast_kernel");
    }
    if (i0 && in1 && C1 == 1 && pad1 == ZR && in2 && C2 == 3 && pad2 == ZR) {
      klaunch(tps_grid_sample_fast_kernel<25, 3, BD, 1, ZR, 3, ZR>, grid, 128, 0, st, theta, to_dev(tps), a, grid_out, B, H, W, chunk);
      return after_launch("tps_grid_sample_fast_kernel");
    }
  }
  if (N == 25 || N == 9) {
    const int chunk = B >= 128 ? kTpsChunk : (B >= 32 ? 8 : 2);
    dim3 grid(cdiv(H * W, 256), cdiv(B, chunk));
    if (N == 25)
      klaunch(tps_grid_sample_batched_kernel<25>, grid, 256, 0, (cudaStream_t)stream, theta, to_dev(tps), a, grid_out, B, H, W, chunk);
    else
      klaunch(tps_grid_sample_batched_kernel<9>, grid, 256, 0, (cudaStream_t)stream, theta, to_dev(tps), a, grid_out, B, H, W, chunk);
    return after_launch("tps_grid_sample_batched_kernel");
  }
  klaunch(tps_grid_sample_kernel, pixel_grid(H * W, B), 256, 0, (cudaStream_t)stream, theta, to_dev(tps), a, grid_out, H, W);
  return after_launch("tps_grid_sample_kernel");
}

int shineon_resample2d_tile(const float* in1, const float* flow, float* out, int B, int C, int H, int W, cudaStream_t stream);  // resample_tile.cu

extern "C" int shineon_resample2d_fwd(const float* in1, const float* flow, float* out, int B, int C, int Hi, int Wi,
                                      int H, int W, int kernel_size, int bilinear, shineon_stream_t stream) {
  SHINEON_REQUIRE(in1 && flow && out, "resample2d_fwd: null pointer");
  SHINEON_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0 && B <= 65535, "resample2d_fwd: bad shape");
  if (kernel_size != 1) return fail(SHINEON_ERR_UNSUPPORTED, "resample2d: kernel_size %d (only 1, as the reference uses)", kernel_size);
  if (Hi != H || Wi != W) return fail(SHINEON_ERR_UNSUPPORTED, "resample2d: input %dx%d != flow %dx%d", Hi, Wi, H, W);
  if (B == 0) return SHINEON_OK;
  if (bilinear && C <= 4 && H >= 32 && W >= 64) {  // shared-memory halo tiles (resample_tile.cu); > 0 = shape does not fit
    const int rc = shineon_resample2d_tile(in1, flow, out, B, C, H, W, (cudaStream_t)stream);
    if (rc <= 0) return rc;
  }
  if (bilinear && C <= 4) {
    constexpr int PPT = 2;
    const dim3 g(cdiv(H * W, 256 * PPT), B);
    cudaStream_t st = (cudaStream_t)stream;
    switch (C) {
      case 1: klaunch(resample2d_fast_kernel<1, PPT>, g, 256, 0, st, in1, flow, out, H, W); break;
      case 2: klaunch(resample2d_fast_kernel<2, PPT>, g, 256, 0, st, in1, flow, out, H, W); break;
      case 3: klaunch(resample2d_fast_kernel<3, PPT>, g, 256, 0, st, in1, flow, out, H, W); break;
      default: klaunch(resample2d_fast_kernel<4, PPT>, g, 256, 0, st, in1, flow, out, H, W); break;
    }
    return after_launch("resample2d_fast_kernel");
  }
  klaunch(resample2d_fwd_kernel, dim3(cdiv(H * W, 256 * kPPT), B), 256, 0, (cudaStream_t)stream, in1, flow, out, C, H, W, bilinear);
  return after_launch("resample2d_fwd_kernel");
}

extern "C" int shineon_resample2d_bwd(const float* in1, const float* flow, const float* grad_out, float* grad_in1,
                                      float* grad_flow, int B, int C, int Hi, int Wi, int H, int W, int kernel_size,
                                      int bilinear, shineon_stream_t stream) {
  SHINEON_REQUIRE(in1 && flow && grad_out && grad_in1 && grad_flow, "resample2d_bwd: null pointer");
  SHINEON_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0 && B <= 65535, "resample2d_bwd: bad shape");
  if (kernel_size != 1) return fail(SHINEON_ERR_UNSUPPORTED, "resample2d: kernel_size %d", kernel_size);
  if (Hi != H || Wi != W) return fail(SHINEON_ERR_UNSUPPORTED, "resample2d: input %dx%d != flow %dx%d", Hi, Wi, H, W);
  (void)bilinear;  // the reference's backward ignores the flag too (resample2d_kernel.cu:76-198)
  if (B == 0) return SHINEON_OK;
  klaunch(resample2d_bwd_in1_kernel, pixel_grid(H * W, B), 256, 0, (cudaStream_t)stream, flow, grad_out, grad_in1, C, H, W);
  int rc = after_launch("resample2d_bwd_in1_kernel");
  if (rc) return rc;
  klaunch(resample2d_bwd_flow_kernel, pixel_grid(H * W, B), 256, 0, (cudaStream_t)stream, in1, flow, grad_out, grad_flow, C, H, W);
  return after_launch("resample2d_bwd_flow_kernel");
}

extern "C" int shineon_channelnorm_fwd(const float* in, float* out, int B, int C, int H, int W, int norm_deg,
                                       shineon_stream_t stream) {
  SHINEON_REQUIRE(in && out, "channelnorm_fwd: null pointer");
  SHINEON_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0, "channelnorm_fwd: bad shape");
  (void)norm_deg;
  if (B == 0) return SHINEON_OK;
  SHINEON_REQUIRE(B <= 65535, "channelnorm_fwd: batch too large");
  const int HW = H * W;
  dim3 grid((HW & 3) == 0 ? cdiv(HW / 4, 256) : min(cdiv(HW, 256), 4096), B);
  klaunch(channelnorm_fwd_kernel, grid, 256, 0, (cudaStream_t)stream, in, out, C, HW);
  return after_launch("channelnorm_fwd_kernel");
}

extern "C" int shineon_channelnorm_bwd(const float* in, const float* out, const float* grad_out, float* grad_in,
                                       int B, int C, int H, int W, int norm_deg, shineon_stream_t stream) {
  SHINEON_REQUIRE(in && out && grad_out && grad_in, "channelnorm_bwd: null pointer");
  SHINEON_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0, "channelnorm_bwd: bad shape");
  (void)norm_deg;
  long total = (long)B * C * H * W;
  if (total == 0) return SHINEON_OK;
  int blocks = (int)((total + 255) / 256 > 65535 ? 65535 : (total + 255) / 256);
  klaunch(channelnorm_bwd_kernel, blocks, 256, 0, (cudaStream_t)stream, in, out, grad_out, grad_in, C, (long)H * W, total);
  return after_launch("channelnorm_bwd_kernel");
}

extern "C" int shineon_tps_warp_u8_planes(const float* theta, const shineon_tps_tables* tps, const unsigned char* cloth_u8,
                                          float* warped, void* z_hi, void* z_lo, int z_cstride, int ctot, int c_off,
                                          int plane_fmt, int B, int H, int W, shineon_stream_t stream) {
  SHINEON_REQUIRE(theta && tps && cloth_u8 && warped && z_hi, "tps_warp_u8_planes: null pointer");
  SHINEON_REQUIRE(tps->Li && tps->P_X && tps->P_Y && tps->grid_X && tps->grid_Y, "tps_warp_u8_planes: null table");
  SHINEON_REQUIRE(plane_fmt == SHINEON_FMT_BF16 || plane_fmt == SHINEON_FMT_FP16, "tps_warp_u8_planes: plane_fmt %d", plane_fmt);
  SHINEON_REQUIRE(B >= 0 && B <= 65535 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "tps_warp_u8_planes: bad shape (H, W even)");
  SHINEON_REQUIRE(ctot >= 3 && c_off >= 0 && c_off + 3 <= ctot && 4 * ctot <= z_cstride, "tps_warp_u8_planes: channel window");
  if (B == 0) return SHINEON_OK;
  const int N = tps->grid_size * tps->grid_size;
  const int chunk = B >= 64 ? kTpsChunk : (B >= 16 ? 8 : 2);
  dim3 grid(cdiv(H * W, 256), cdiv(B, chunk));
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 25)
    klaunch(tps_warp_u8_planes_kernel<25>, grid, 128, 0, st, theta, to_dev(tps), cloth_u8, warped, (plane_t*)z_hi, (plane_t*)z_lo,
                                                         z_cstride, ctot, c_off, plane_fmt, B, H, W, chunk);
  else if (N == 9)
    klaunch(tps_warp_u8_planes_kernel<9>, grid, 128, 0, st, theta, to_dev(tps), cloth_u8, warped, (plane_t*)z_hi, (plane_t*)z_lo,
                                                        z_cstride, ctot, c_off, plane_fmt, B, H, W, chunk);
  else
    return fail(SHINEON_ERR_UNSUPPORTED, "tps_warp_u8_planes: grid_size %d (3 and 5 are compiled)", tps->grid_size);
  return after_launch("tps_warp_u8_planes_kernel");
}
