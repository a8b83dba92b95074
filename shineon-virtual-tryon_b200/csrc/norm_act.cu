// Memory-bound passes of the U-Net / TOM path, all NHWC:
//   * nchw_to_planes   : reference-layout input -> activated bf16 hi/lo planes (unet.py:132 on the block input)
//   * instnorm_act     : nn.InstanceNorm2d(affine=False) (unet.py:133,135) + following activation
//   * upsample2x_cat   : up_act -> torch.cat([x, x'],1) -> nn.Upsample(2,"bilinear") (unet.py:138,155,166,198)
//   * tom_compose      : tanh / sigmoid / mask compose of UnetMaskModel.forward (unet_mask_model.py:74-135)
#include "common.cuh"

namespace shineon {

// ------------------------------------------------------------------------------ nchw -> planes
__global__ void __launch_bounds__(256)
    nchw_to_planes_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1,
                          plane_t* __restrict__ yh, plane_t* __restrict__ yl, int HW, int cpad,
                          int act, float act_param, int fmt) {
  pdl_grid_sync();
  const int n = blockIdx.y;
  const int groups = (C0 + C1 + 7) / 8;
  const long total = (long)HW * groups;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int p = (int)(e % HW);  // pixel fastest: coalesced channel-plane reads
    const int g = (int)(e / HW);
    __align__(16) plane_t hi[8];
    __align__(16) plane_t lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      float v = 0.f;
      if (c < C0)
        v = x0[((long)n * C0 + c) * HW + p];
      else if (c < C0 + C1)
        v = x1[((long)n * C1 + (c - C0)) * HW + p];
      v = (c < C0 + C1) ? apply_act(v, act, act_param) : 0.f;
      split16(v, fmt, hi[j], lo[j]);
    }
    const long o = ((long)n * HW + p) * cpad + g * 8;
    *reinterpret_cast<uint4*>(yh + o) = *reinterpret_cast<const uint4*>(hi);
    if (yl) *reinterpret_cast<uint4*>(yl + o) = *reinterpret_cast<const uint4*>(lo);
  }
}

// ------------------------------------------------------------------------------ planes -> nchw f32
// (feeds ops with the reference's NCHW interface, e.g. the FlowNetC correlation, from a tensor-core layer output)
__global__ void __launch_bounds__(256)
    planes_to_nchw_kernel(const plane_t* __restrict__ xh, const plane_t* __restrict__ xl, int cstride,
                          float* __restrict__ y, int HW, int C, int fmt) {
  pdl_grid_sync();
  __shared__ float tile[32][33];
  const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {  // read: channel fastest
    const int p = p0 + i, c = c0 + tx;
    float v = 0.f;
    if (p < HW && c < C) {
      const long o = ((long)n * HW + p) * cstride + c;
      v = xl ? join16(xh[o], xl[o], fmt) : load16(xh[o], fmt);
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {  // write: pixel fastest
    const int c = c0 + i, p = p0 + tx;
    if (p < HW && c < C) y[((long)n * C + c) * HW + p] = tile[tx][i];
  }
}

// ------------------------------------------------------------------------------ nchw -> im2col planes
// First layers have tiny Cin (3 / 10 / 22): one K-block per filter tap would waste >= 2/3 of every TMA box and
// MMA on zero padding.  Instead the layout conversion writes, per OUTPUT pixel, the K-vector
//   k = (fy*kw + fx)*C + c   (tap-major, channel-minor; zero for out-of-image taps and k >= kh*kw*C)
// so the convolution becomes a 1x1 GEMM with K = pad64(kh*kw*C).
template <int FMT>
__global__ void __launch_bounds__(256)
    nchw_im2col_planes_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1,
                              plane_t* __restrict__ yh, plane_t* __restrict__ yl, int H, int W, int kh, int kw,
                              int stride, int pad, int Ho, int Wo, int kpad, int act, float act_param) {
  pdl_grid_sync();
  // grid: x = (pair of k-groups, output column) tiles, y = output row, z = image; output column fastest so
  // neighbouring threads read neighbouring input columns of the same NCHW channel plane.  Each thread produces 16
  // consecutive k (two 16-byte chunks per plane): 16 independent gathers in flight.
  const int n = blockIdx.z, oh = blockIdx.y;
  const int pairs = kpad >> 4;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= Wo * pairs) return;
  const int gp = t / Wo, ow = t - gp * Wo;
  const int C = C0 + C1;
  const int K = kh * kw * C;
  const int HW = H * W;
  int k = gp * 16;
  int tap = k / C, c = k - tap * C;
  int fy = tap / kw, fx = tap - fy * kw;
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j, ++k) {
    float val = 0.f;
    if (k < K) {
      const int iy = oh * stride + fy - pad, ix = ow * stride + fx - pad;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W)
        val = c < C0 ? __ldg(x0 + ((long)n * C0 + c) * HW + iy * W + ix)
                     : __ldg(x1 + ((long)n * C1 + (c - C0)) * HW + iy * W + ix);
    }
    v[j] = val;
    if (++c == C) {  // next tap
      c = 0;
      if (++fx == kw) { fx = 0; ++fy; }
    }
  }
  __align__(16) plane_t hi[16];
  __align__(16) plane_t lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) split16(apply_act(v[j], act, act_param), FMT, hi[j], lo[j]);
  const long o = (((long)n * Ho + oh) * Wo + ow) * kpad + gp * 16;
  reinterpret_cast<uint4*>(yh + o)[0] = reinterpret_cast<const uint4*>(hi)[0];
  reinterpret_cast<uint4*>(yh + o)[1] = reinterpret_cast<const uint4*>(hi)[1];
  if (yl) {
    reinterpret_cast<uint4*>(yl + o)[0] = reinterpret_cast<const uint4*>(lo)[0];
    reinterpret_cast<uint4*>(yl + o)[1] = reinterpret_cast<const uint4*>(lo)[1];
  }
}

// (A tiled variant -- input patch staged in shared memory, consecutive threads on consecutive 16-byte chunks of a pixel's K
// vector so that a warp stores 512 contiguous bytes per plane -- was measured in r02 on the six FlowNet2 stems at batch 16:
// 1.16 ms against 0.96 ms for this gather form; the pass is bound by the 1.9 GB it writes, not by store coalescing.)

// ------------------------------------------------------------------------------ nchw -> shifted space-to-depth planes
// A 4x4 / stride 2 / pad 1 conv over x equals a 2x2 / stride 1 / pad 0 conv over
//   z[Y][X][(py*2+px)*C + c] = x[c][2Y-1+py][2X-1+px]   (zero outside the image),   Y in [0, H/2], X in [0, W/2]
// (filter row fy = 2*ay + py).  Unlike the im2col form (16 taps per output pixel = 4x the input) z holds every input
// element exactly once: the layout conversion writes 4x fewer bytes and converts each element to 16-bit once.
template <int FMT>
__global__ void __launch_bounds__(256)
    nchw_s2d_planes_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1, int C1,
                           plane_t* __restrict__ yh, plane_t* __restrict__ yl, int H, int W, int cpad) {
  pdl_grid_sync();
  const int n = blockIdx.z, Y = blockIdx.y;
  const int Wz = W / 2 + 1;
  const int groups = cpad >> 3;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= Wz * groups) return;
  const int g = t / Wz, X = t - g * Wz;  // X fastest: neighbouring threads read neighbouring columns of one channel plane
  const int C = C0 + C1;
  const int HW = H * W;
  __align__(16) plane_t hi[8];
  __align__(16) plane_t lo[8];
  int k = g * 8;
  int q = k / C, c = k - q * C;  // q = py*2+px
#pragma unroll
  for (int j = 0; j < 8; ++j, ++k) {
    float v = 0.f;
    if (q < 4) {
      const int iy = 2 * Y - 1 + (q >> 1), ix = 2 * X - 1 + (q & 1);
      if (iy >= 0 && iy < H && ix >= 0 && ix < W)
        v = c < C0 ? __ldg(x0 + ((long)n * C0 + c) * HW + iy * W + ix)
                   : __ldg(x1 + ((long)n * C1 + (c - C0)) * HW + iy * W + ix);
    }
    split16(v, FMT, hi[j], lo[j]);
    if (++c == C) { c = 0; ++q; }
  }
  const long o = (((long)n * (H / 2 + 1) + Y) * Wz + X) * cpad + g * 8;
  *reinterpret_cast<uint4*>(yh + o) = *reinterpret_cast<const uint4*>(hi);
  if (yl) *reinterpret_cast<uint4*>(yl + o) = *reinterpret_cast<const uint4*>(lo);
}

// (A shared-memory tiled variant -- coalesced 256-byte channel-row reads, whole-pixel writes -- was measured SLOWER than
// this direct form on B200: 0.41 vs 0.35 ms for 22 channels, 0.28 vs 0.16 ms for 10 at 80 frames; the direct form's
// strided accesses are absorbed by L1/L2 and it has no barrier.  profiles/r01_memory_ops.md.)

// ------------------------------------------------------------------------------ tap-stacked 3x3 conv: col2im
// For a 3x3 conv with very few output channels (the U-Net's final 128 -> 4 layer) the implicit GEMM is run
// "transposed": one 1x1 GEMM produces, for every INPUT pixel q, the 9*Cout partial products
//   t[q, (fy*3+fx)*Cout + co] = <x[q,:], w[co,:,fy,fx]>
// (the activation is read once instead of once per tap) and this kernel sums the shifted partials:
//   out[n,oh,ow,co] = bias[co] + sum_{fy,fx} t[n, oh+fy-1, ow+fx-1, (fy*3+fx)*Cout + co]   (zero outside the image)
__global__ void __launch_bounds__(256)
    col2im3x3_kernel(const float* __restrict__ t, const float* __restrict__ bias, float* __restrict__ y, int H, int W,
                     int Cout, int tstride) {
  pdl_grid_sync();
  const int n = blockIdx.y;
  const long total = (long)H * W * Cout;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int co = (int)(e % Cout);
    const long p = e / Cout;
    const int ow = (int)(p % W), oh = (int)(p / W);
    float acc = bias ? __ldg(bias + co) : 0.f;
#pragma unroll
    for (int fy = 0; fy < 3; ++fy) {
      const int iy = oh + fy - 1;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int fx = 0; fx < 3; ++fx) {
        const int ix = ow + fx - 1;
        if (ix < 0 || ix >= W) continue;
        acc += __ldg(t + (((long)n * H + iy) * W + ix) * tstride + (fy * 3 + fx) * Cout + co);
      }
    }
    y[((long)n * H * W + p) * Cout + co] = acc;
  }
}

// ------------------------------------------------------------------------------ bilinear x2 -> 3x3 conv, at low resolution
// nn.Upsample(x2, bilinear, align_corners=False) followed by Conv2d(3x3, pad 1) (unet.py:138-146,155,166) is linear in
// the low-resolution tensor, and the channel contraction commutes with the spatial interpolation:
//   conv(up(x))[n,r,s,co] = bias[co] + sum_{fy,fx} [0 <= r+fy-1 < 2h, 0 <= s+fx-1 < 2w] * up(t_{fy,fx,co})[n, r+fy-1, s+fx-1]
// with t_{fy,fx,co}[n,i,j] = <x[n,i,j,:], w[co,:,fy,fx]> -- the tap-stacked 1x1 GEMM of the LOW-resolution tensor
// (4x fewer pixels => 4x fewer tensor-core FLOPs, and the upsampled activation is never written or read).
// This kernel evaluates the right-hand side: one thread owns low-res pixel (i,j) x VEC channels and produces its 2x2
// output pixels.  Row r = 2i+py+fy-1 of up(.) interpolates low-res rows {i-1,i,i+1} with one of four coefficient
// triples (clamped at the border like PyTorch's upsample, zero where the conv's zero padding applies):
//   A: r = 2i-1  (0.75, 0.25, 0)  [zero when i == 0]        B: r = 2i    (0.25, 0.75, 0)  [(0,1,0) when i == 0]
//   C: r = 2i+1  (0, 0.75, 0.25)  [(0,1,0) when i == h-1]   D: r = 2i+2  (0, 0.25, 0.75)  [zero when i == h-1]
// (fy,py) -> triple: (0,0) A, (0,1) B, (1,0) B, (1,1) C, (2,0) C, (2,1) D; same for columns.
struct Up3 { float c[4][3]; };  // A, B, C, D
__device__ __forceinline__ Up3 up3_coeffs(int i, int n) {
  Up3 u;
  const bool first = i == 0, last = i == n - 1;
  u.c[0][0] = first ? 0.f : 0.75f; u.c[0][1] = first ? 0.f : 0.25f; u.c[0][2] = 0.f;
  u.c[1][0] = first ? 0.f : 0.25f; u.c[1][1] = first ? 1.f : 0.75f; u.c[1][2] = 0.f;
  u.c[2][0] = 0.f; u.c[2][1] = last ? 1.f : 0.75f; u.c[2][2] = last ? 0.f : 0.25f;
  u.c[3][0] = 0.f; u.c[3][1] = last ? 0.f : 0.25f; u.c[3][2] = last ? 0.f : 0.75f;
  return u;
}
// structural support of triple (f + p) in {A,B,C,D}: A,B touch rows {0,1}; C,D rows {1,2}
__host__ __device__ constexpr bool up3_has(int f, int p, int a) { return (f + p) < 2 ? a < 2 : a > 0; }

template <int VEC>
__global__ void __launch_bounds__(256)
    upconv3x3_gather_kernel(const float* __restrict__ t, const float* __restrict__ bias, float* __restrict__ y, int h,
                            int w, int Cout, int tstride, int gpb, int tile_w, int tiles_x) {
  pdl_grid_sync();
  const int n = blockIdx.z;
  const int g = blockIdx.y * gpb + (int)threadIdx.x % gpb;
  const int pix = (int)threadIdx.x / gpb;
  const int tile_h = (256 / gpb) / tile_w;
  const int i = ((int)blockIdx.x / tiles_x) * tile_h + pix / tile_w;
  const int j = ((int)blockIdx.x % tiles_x) * tile_w + pix % tile_w;
  const int co = g * VEC;
  if (i >= h || j >= w || co >= Cout) return;
  const Up3 cy = up3_coeffs(i, h), cx = up3_coeffs(j, w);
  const int rows[3] = {max(i - 1, 0), i, min(i + 1, h - 1)};
  const int cols[3] = {max(j - 1, 0), j, min(j + 1, w - 1)};
  float out[2][2][VEC];
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int k = 0; k < VEC; ++k) out[p][q][k] = bias ? __ldg(bias + co + k) : 0.f;
  const float* tn = t + (long)n * h * w * tstride + co;
#pragma unroll
  for (int fy = 0; fy < 3; ++fy) {
#pragma unroll
    for (int fx = 0; fx < 3; ++fx) {
      const float* tt = tn + (fy * 3 + fx) * Cout;
      // rows needed by this tap: fy=0 {0,1}, fy=1 {0,1,2}, fy=2 {1,2}; same for columns
      float v[3][3][VEC];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (!(up3_has(fy, 0, a) || up3_has(fy, 1, a))) continue;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          if (!(up3_has(fx, 0, b) || up3_has(fx, 1, b))) continue;
          const float* src = tt + ((long)rows[a] * w + cols[b]) * tstride;
          if constexpr (VEC == 4) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(src));
            v[a][b][0] = f.x; v[a][b][1] = f.y; v[a][b][2] = f.z; v[a][b][3] = f.w;
          } else {
            v[a][b][0] = __ldg(src);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (!(up3_has(fy, 0, a) || up3_has(fy, 1, a))) continue;
        float tmp[2][VEC];  // column-interpolated row a, for the two output columns
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            float s = 0.f;
#pragma unroll
            for (int b = 0; b < 3; ++b)
              if (up3_has(fx, q, b)) s = fmaf(cx.c[fx + q][b], v[a][b][k], s);
            tmp[q][k] = s;
          }
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          if (!up3_has(fy, p, a)) continue;
          const float wy = cy.c[fy + p][a];
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int k = 0; k < VEC; ++k) out[p][q][k] = fmaf(wy, tmp[q][k], out[p][q][k]);
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float* dst = y + (((long)n * 2 * h + 2 * i + p) * (2 * w) + 2 * j + q) * Cout + co;
      if constexpr (VEC == 4)
        *reinterpret_cast<float4*>(dst) = make_float4(out[p][q][0], out[p][q][1], out[p][q][2], out[p][q][3]);
      else
        dst[0] = out[p][q][0];
    }
}

// ------------------------------------------------------------------------------ instance norm
// Pass 1: per-(n,c) sum and sum of squares.  fp32 partials per CTA, fp64 atomics across CTAs
// (E[x^2]-E[x]^2 is then evaluated in fp64, so cancellation is not an issue).
template <int CB>  // channels per CTA (power of two <= 32)
__global__ void __launch_bounds__(256)
    instnorm_stats_kernel(const float* __restrict__ x, double* __restrict__ ws, int HW, int C, int pix_per_cta) {
  pdl_grid_sync();
  constexpr int PP = 256 / CB;
  __shared__ float s_sum[256], s_sq[256];
  const int n = blockIdx.z;
  const int c = blockIdx.y * CB + (threadIdx.x % CB);
  const int pl = threadIdx.x / CB;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(p0 + pix_per_cta, HW);
  float s = 0.f, q = 0.f;
  if (c < C) {
    const float* base = x + (long)n * HW * C + c;
#pragma unroll 4
    for (int p = p0 + pl; p < p1; p += PP) {
      float v = __ldg(base + (long)p * C);
      s += v;
      q = fmaf(v, v, q);
    }
  }
  s_sum[threadIdx.x] = s;
  s_sq[threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.x < CB && c < C) {
    float ts = 0.f, tq = 0.f;
    for (int i = 0; i < PP; ++i) {
      ts += s_sum[i * CB + threadIdx.x];
      tq += s_sq[i * CB + threadIdx.x];
    }
    atomicAdd(ws + ((long)n * C + c) * 2 + 0, (double)ts);
    atomicAdd(ws + ((long)n * C + c) * 2 + 1, (double)tq);
  }
}

// Pass 2: y = act((x - mean) * rsqrt(var + eps)); VEC = 8 / 4 / 1 consecutive channels per thread (largest that divides
// C): 8 channels = two 128-bit loads in flight per thread and one 128-bit store per plane.
template <int FMT, int ACT, int VEC>  // ACT < 0: runtime activation id
__global__ void __launch_bounds__(256)
    instnorm_apply_kernel(const float* __restrict__ x, const double* __restrict__ ws, float* __restrict__ yf,
                          plane_t* __restrict__ yh, plane_t* __restrict__ yl, int HW, int C, int cpad,
                          float eps, int do_norm, int act_rt, float act_param) {
  pdl_grid_sync();
  extern __shared__ float s_tab[];  // mean[C], rstd[C]
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean = 0.f, rstd = 1.f;
    if (do_norm) {
      double s = ws[((long)n * C + c) * 2], q = ws[((long)n * C + c) * 2 + 1];
      double m = s / HW;
      double var = q / HW - m * m;  // biased variance (F.instance_norm)
      if (var < 0.0) var = 0.0;
      mean = (float)m;
      rstd = (float)(1.0 / sqrt(var + (double)eps));
    }
    s_tab[c] = mean;
    s_tab[C + c] = rstd;
  }
  __syncthreads();
  const int cg = C / VEC;
  const unsigned total = (unsigned)HW * (unsigned)cg;
  const int act = ACT < 0 ? act_rt : ACT;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const unsigned p = e / (unsigned)cg;
    const int g = (int)(e - p * (unsigned)cg);
    const long xi = ((long)n * HW + p) * C + g * VEC;
    float v[VEC];
    if constexpr (VEC >= 4) {
#pragma unroll
      for (int j = 0; j < VEC; j += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(x + xi + j));
        v[j] = t.x; v[j + 1] = t.y; v[j + 2] = t.z; v[j + 3] = t.w;
      }
    } else {
      v[0] = x[xi];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = g * VEC + j;
      v[j] = apply_act((v[j] - s_tab[c]) * s_tab[C + c], act, act_param);
    }
    if (yf) {
      if constexpr (VEC >= 4) {
#pragma unroll
        for (int j = 0; j < VEC; j += 4)
          *reinterpret_cast<float4*>(yf + xi + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
        yf[xi] = v[0];
      }
    }
    if (yh) {
      const long po = ((long)n * HW + p) * cpad + g * VEC;
      if constexpr (VEC >= 4) {
        uint32_t h[VEC / 2], l[VEC / 2];
#pragma unroll
        for (int j = 0; j < VEC / 2; ++j) split16x2(v[2 * j], v[2 * j + 1], FMT, h[j], l[j]);
        if constexpr (VEC == 8) {
          *reinterpret_cast<uint4*>(yh + po) = make_uint4(h[0], h[1], h[2], h[3]);
          if (yl) *reinterpret_cast<uint4*>(yl + po) = make_uint4(l[0], l[1], l[2], l[3]);
        } else {
          *reinterpret_cast<uint2*>(yh + po) = make_uint2(h[0], h[1]);
          if (yl) *reinterpret_cast<uint2*>(yl + po) = make_uint2(l[0], l[1]);
        }
      } else {
        plane_t h, l;
        split16(v[0], FMT, h, l);
        yh[po] = h;
        if (yl) yl[po] = l;
      }
    }
  }
}

// ------------------------------------------------------------------------------ upsample x2 + concat
// PyTorch upsample_bilinear2d, align_corners=False, scale 2: src = max(0.5*(dst+0.5)-0.5, 0).
__device__ __forceinline__ void up2_index(int d, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = 0.5f * ((float)d + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

template <int FMT, int ACT>
__device__ __forceinline__ void load8(const plane_t* __restrict__ h, const plane_t* __restrict__ l, long off,
                                      int act_rt, float act_param, float (&v)[8]) {
  const int act = ACT < 0 ? act_rt : ACT;
  uint4 uh = __ldg(reinterpret_cast<const uint4*>(h + off));
  const plane_t* ph = reinterpret_cast<const plane_t*>(&uh);
  if (l) {
    uint4 ul = __ldg(reinterpret_cast<const uint4*>(l + off));
    const plane_t* pl = reinterpret_cast<const plane_t*>(&ul);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = apply_act(join16(ph[j], pl[j], FMT), act, act_param);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = apply_act(load16(ph[j], FMT), act, act_param);
  }
}

// One thread = one 2x2 OUTPUT block x 8 channels.  For scale 2 / align_corners=False the outputs (2i+1, 2i+2) x
// (2j+1, 2j+2) read exactly the source pixels (i, i+1) x (j, j+1): four source loads (hi+lo joined once) feed four
// outputs, instead of four loads per output.  Blocks i = -1 and i = H-1 are the clamped borders (one valid output
// row).  grid: x = (block column, 8-channel group) tiles, y = block row (H+1), z = image.
template <int FMT, int ACT>  // ACT < 0: runtime activation id
__global__ void __launch_bounds__(256)
    upsample2x_cat_kernel(const plane_t* __restrict__ s0h, const plane_t* __restrict__ s0l, int c0pad,
                          const plane_t* __restrict__ s1h, const plane_t* __restrict__ s1l, int c1pad,
                          plane_t* __restrict__ yh, plane_t* __restrict__ yl, int H, int W, int act,
                          float act_param) {
  pdl_grid_sync();
  const int n = blockIdx.z, bi = (int)blockIdx.y - 1;
  const int ctot = c0pad + c1pad;
  const int groups = ctot >> 3;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= (W + 1) * groups) return;
  const int bjp = t / groups, g = t - bjp * groups;
  const int bj = bjp - 1;
  const int c = g * 8;
  const plane_t *sh, *sl;
  int cp, cc;
  if (c < c0pad) { sh = s0h; sl = s0l; cp = c0pad; cc = c; } else { sh = s1h; sl = s1l; cp = c1pad; cc = c - c0pad; }
  const int r0 = max(bi, 0), r1 = min(bi + 1, H - 1), q0 = max(bj, 0), q1 = min(bj + 1, W - 1);
  const long rb = (long)n * H * W;
  float a00[8], a01[8], a10[8], a11[8];
  load8<FMT, ACT>(sh, sl, (rb + r0 * W + q0) * cp + cc, act, act_param, a00);
  load8<FMT, ACT>(sh, sl, (rb + r0 * W + q1) * cp + cc, act, act_param, a01);
  load8<FMT, ACT>(sh, sl, (rb + r1 * W + q0) * cp + cc, act, act_param, a10);
  load8<FMT, ACT>(sh, sl, (rb + r1 * W + q1) * cp + cc, act, act_param, a11);
  const int Wo = 2 * W, Ho = 2 * H;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
    const int oy = 2 * bi + 1 + dy;
    if (oy < 0 || oy >= Ho) continue;
    int i0, i1;
    float ly0, ly1;
    up2_index(oy, H, i0, i1, ly0, ly1);  // (i0, i1) == (r0, r1) up to the clamped borders, where the weight of the other row is 0
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const int ox = 2 * bj + 1 + dx;
      if (ox < 0 || ox >= Wo) continue;
      int j0, j1;
      float lx0, lx1;
      up2_index(ox, W, j0, j1, lx0, lx1);
      __align__(16) plane_t hi[8];
      __align__(16) plane_t lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // ATen: h0lambda*(w0lambda*p00 + w1lambda*p01) + h1lambda*(w0lambda*p10 + w1lambda*p11)
        const float v = ly0 * (lx0 * a00[j] + lx1 * a01[j]) + ly1 * (lx0 * a10[j] + lx1 * a11[j]);
        split16(v, FMT, hi[j], lo[j]);
      }
      const long o = (((long)n * Ho + oy) * Wo + ox) * ctot + c;
      *reinterpret_cast<uint4*>(yh + o) = *reinterpret_cast<const uint4*>(hi);
      if (yl) *reinterpret_cast<uint4*>(yl + o) = *reinterpret_cast<const uint4*>(lo);
    }
  }
}

// ------------------------------------------------------------------------------ 8-bit image writer
// visualization.save_images (visualization.py:73-76): tensor = (img + 1) * 0.5 * 255 -> clamp(0, 255) ->
// numpy astype("uint8") (truncation) -> channel-last.  Separate roundings like the reference's three torch ops (no FMA).
__device__ __forceinline__ uint8_t image_u8(float x) {
  float t = __fmul_rn(__fmul_rn(__fadd_rn(x, 1.f), 0.5f), 255.f);
  t = fminf(fmaxf(t, 0.f), 255.f);
  return (uint8_t)(int)t;
}

// x f32 [B,C,H,W] (C = 1 or 3) -> y uint8 [B,H,W,C]; 4 consecutive pixels per thread (float4 loads per channel plane,
// 32-bit stores of the interleaved bytes).
template <int C>
__global__ void __launch_bounds__(256)
    image_to_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ y, int HW) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  const int p0 = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (p0 >= HW) return;
  uint8_t v[4 * C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(x + ((long)b * C + c) * HW + p0));
    v[c] = image_u8(f.x); v[C + c] = image_u8(f.y); v[2 * C + c] = image_u8(f.z); v[3 * C + c] = image_u8(f.w);
  }
  uint32_t* dst = reinterpret_cast<uint32_t*>(y + ((long)b * HW + p0) * C);
#pragma unroll
  for (int i = 0; i < C; ++i)
    dst[i] = (uint32_t)v[4 * i] | ((uint32_t)v[4 * i + 1] << 8) | ((uint32_t)v[4 * i + 2] << 16) | ((uint32_t)v[4 * i + 3] << 24);
}

__global__ void __launch_bounds__(256)
    image_to_u8_scalar_kernel(const float* __restrict__ x, uint8_t* __restrict__ y, int HW, int C) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  for (long e = (long)blockIdx.x * 256 + threadIdx.x; e < (long)HW * C; e += (long)gridDim.x * 256) {
    const int c = (int)(e % C);
    const long p = e / C;
    y[(long)b * HW * C + e] = image_u8(x[((long)b * C + c) * HW + p]);
  }
}

// ------------------------------------------------------------------------------ TOM compose
// One thread = 4 consecutive pixels of one image: the NHWC U-Net output is read as whole pixels, every NCHW plane is
// written with float4 stores, the optional 8-bit try-on image with three 32-bit stores.
template <bool VEC4>
__global__ void __launch_bounds__(256)
    tom_compose_kernel(const float* __restrict__ u, int Cout, const float* __restrict__ cloth,
                       const float* __restrict__ warped_prev, float* __restrict__ p_rend, float* __restrict__ masks,
                       float* __restrict__ p_tryon, float* __restrict__ fmasks, uint8_t* __restrict__ tryon_u8, int HW,
                       int nf, int f, int flow_warp) {
  pdl_grid_sync();
  constexpr int PX = VEC4 ? 4 : 1;
  const int b = blockIdx.y;
  for (int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * PX; p0 < HW; p0 += gridDim.x * blockDim.x * PX) {
    float m[PX], fm[PX], r[3][PX];
#pragma unroll
    for (int i = 0; i < PX; ++i) {
      const float* up = u + ((long)b * HW + p0 + i) * Cout;
      m[i] = 1.f / (1.f + expf(-__ldg(up + 3 * nf + f)));  // F.sigmoid (unet_mask_model.py:85)
      fm[i] = flow_warp ? 1.f / (1.f + expf(-__ldg(up + 4 * nf + f))) : 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) r[k][i] = tanhf(__ldg(up + 3 * f + k));  // F.tanh (unet_mask_model.py:84)
    }
    auto store = [&](float* base, long o, const float (&v)[PX]) {
      if (!base) return;
      if constexpr (VEC4) *reinterpret_cast<float4*>(base + o) = make_float4(v[0], v[1], v[2], v[3]);
      else base[o] = v[0];
    };
    auto load = [&](const float* base, long o, float (&v)[PX]) {
      if constexpr (VEC4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(base + o));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
        v[0] = __ldg(base + o);
      }
    };
    store(masks, ((long)b * nf + f) * HW + p0, m);
    if (flow_warp) store(fmasks, ((long)b * nf + f) * HW + p0, fm);
    uint8_t q[3][PX];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const long o = ((long)b * 3 * nf + 3 * f + k) * HW + p0;
      store(p_rend, o, r[k]);
      float rr[PX], c[PX], t[PX];
#pragma unroll
      for (int i = 0; i < PX; ++i) rr[i] = r[k][i];
      if (warped_prev) {  // unet_mask_model.py:118-121
        float w[PX];
        load(warped_prev, ((long)b * 3 + k) * HW + p0, w);
#pragma unroll
        for (int i = 0; i < PX; ++i) rr[i] = (1.f - fm[i]) * w[i] + fm[i] * r[k][i];
      }
      load(cloth, o, c);
#pragma unroll
      for (int i = 0; i < PX; ++i) {
        t[i] = (1.f - m[i]) * rr[i] + m[i] * c[i];  // unet_mask_model.py:126-129
        q[k][i] = image_u8(t[i]);
      }
      store(p_tryon, o, t);
    }
    if (tryon_u8) {  // [B, nf, H, W, 3]
      uint8_t* dst = tryon_u8 + (((long)b * nf + f) * HW + p0) * 3;
      if constexpr (VEC4) {
        uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
        d32[0] = (uint32_t)q[0][0] | ((uint32_t)q[1][0] << 8) | ((uint32_t)q[2][0] << 16) | ((uint32_t)q[0][1] << 24);
        d32[1] = (uint32_t)q[1][1] | ((uint32_t)q[2][1] << 8) | ((uint32_t)q[0][2] << 16) | ((uint32_t)q[1][2] << 24);
        d32[2] = (uint32_t)q[2][2] | ((uint32_t)q[0][3] << 8) | ((uint32_t)q[1][3] << 16) | ((uint32_t)q[2][3] << 24);
      } else {
        dst[0] = q[0][0]; dst[1] = q[1][0]; dst[2] = q[2][0];
      }
    }
  }
}

static inline int grid_x(long total, int threads) {
  long b = (total + threads - 1) / threads;
  return (int)(b > 148 * 32 ? 148 * 32 : (b < 1 ? 1 : b));
}

}  // namespace shineon

using namespace shineon;

int launch_instnorm_stats(const float* x, double* ws, int N, int HW, int C, cudaStream_t stream);

#define SHINEON_REQUIRE_FMT(f, who) SHINEON_REQUIRE((f) == SHINEON_FMT_BF16 || (f) == SHINEON_FMT_FP16, who ": plane_fmt %d", (f))

extern "C" int shineon_nchw_to_planes(const float* x0, int C0, const float* x1, int C1, void* y_hi, void* y_lo,
                                      int N, int H, int W, int cpad, int act, float act_param, int plane_fmt,
                                      shineon_stream_t stream) {
  SHINEON_REQUIRE_FMT(plane_fmt, "nchw_to_planes");
  SHINEON_REQUIRE(x0 && y_hi && C0 > 0, "nchw_to_planes: null pointer");
  SHINEON_REQUIRE((x1 == nullptr) == (C1 == 0), "nchw_to_planes: x1/C1 mismatch");
  SHINEON_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0, "nchw_to_planes: bad shape");
  SHINEON_REQUIRE(cpad % 8 == 0 && cpad >= C0 + C1, "nchw_to_planes: cpad %d too small / not a multiple of 8", cpad);
  const int HW = H * W;
  dim3 grid(grid_x((long)HW * ((C0 + C1 + 7) / 8), 256), N);
  klaunch(nchw_to_planes_kernel, grid, 256, 0, (cudaStream_t)stream, x0, C0, x1, C1, (plane_t*)y_hi, (plane_t*)y_lo,
                                                               HW, cpad, act, act_param, plane_fmt);
  return after_launch("nchw_to_planes_kernel");
}

extern "C" int shineon_planes_to_nchw(const void* x_hi, const void* x_lo, int x_cstride, float* y, int N, int H, int W,
                                      int C, int plane_fmt, shineon_stream_t stream) {
  SHINEON_REQUIRE_FMT(plane_fmt, "planes_to_nchw");
  SHINEON_REQUIRE(x_hi && y && N > 0 && N <= 65535 && H > 0 && W > 0 && C > 0 && x_cstride >= C, "planes_to_nchw: bad argument");
  dim3 grid(cdiv(H * W, 32), cdiv(C, 32), N);
  klaunch(planes_to_nchw_kernel, grid, 256, 0, (cudaStream_t)stream, (const plane_t*)x_hi, (const plane_t*)x_lo, x_cstride, y,
                                                               H * W, C, plane_fmt);
  return after_launch("planes_to_nchw_kernel");
}

extern "C" int shineon_nchw_im2col_planes(const float* x0, int C0, const float* x1, int C1, void* y_hi, void* y_lo,
                                         int N, int H, int W, int kh, int kw, int stride, int pad, int Ho, int Wo,
                                         int kpad, int act, float act_param, int plane_fmt, shineon_stream_t stream) {
  SHINEON_REQUIRE_FMT(plane_fmt, "nchw_im2col_planes");
  SHINEON_REQUIRE(x0 && y_hi && C0 > 0, "nchw_im2col_planes: null pointer");
  SHINEON_REQUIRE((x1 == nullptr) == (C1 == 0), "nchw_im2col_planes: x1/C1 mismatch");
  SHINEON_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0, "nchw_im2col_planes: bad shape");
  SHINEON_REQUIRE(Ho == (H + 2 * pad - kh) / stride + 1 && Wo == (W + 2 * pad - kw) / stride + 1, "nchw_im2col_planes: Ho/Wo");
  SHINEON_REQUIRE(kpad % 16 == 0 && kpad >= kh * kw * (C0 + C1), "nchw_im2col_planes: kpad %d too small / not a multiple of 16", kpad);
  SHINEON_REQUIRE(Ho <= 65535, "nchw_im2col_planes: Ho too large");
  dim3 grid(cdiv(Wo * (kpad / 16), 256), Ho, N);
  if (plane_fmt == SHINEON_FMT_FP16)
    klaunch(nchw_im2col_planes_kernel<SHINEON_FMT_FP16>, grid, 256, 0, (cudaStream_t)stream, 
        x0, C0, x1, C1, (plane_t*)y_hi, (plane_t*)y_lo, H, W, kh, kw, stride, pad, Ho, Wo, kpad, act, act_param);
  else
    klaunch(nchw_im2col_planes_kernel<SHINEON_FMT_BF16>, grid, 256, 0, (cudaStream_t)stream, 
        x0, C0, x1, C1, (plane_t*)y_hi, (plane_t*)y_lo, H, W, kh, kw, stride, pad, Ho, Wo, kpad, act, act_param);
  return after_launch("nchw_im2col_planes_kernel");
}

extern "C" int shineon_nchw_s2d_planes(const float* x0, int C0, const float* x1, int C1, void* y_hi, void* y_lo, int N, int H,
                                       int W, int cpad, int plane_fmt, shineon_stream_t stream) {
  SHINEON_REQUIRE_FMT(plane_fmt, "nchw_s2d_planes");
  SHINEON_REQUIRE(x0 && y_hi && C0 > 0 && (x1 == nullptr) == (C1 == 0), "nchw_s2d_planes: bad input tensors");
  SHINEON_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && H / 2 + 1 <= 65535, "nchw_s2d_planes: bad shape (H, W must be even)");
  SHINEON_REQUIRE(cpad % 8 == 0 && cpad >= 4 * (C0 + C1), "nchw_s2d_planes: cpad %d too small / not a multiple of 8", cpad);
  dim3 grid(cdiv((W / 2 + 1) * (cpad / 8), 256), H / 2 + 1, N);
  if (plane_fmt == SHINEON_FMT_FP16)
    klaunch(nchw_s2d_planes_kernel<SHINEON_FMT_FP16>, grid, 256, 0, (cudaStream_t)stream, x0, C0, x1, C1, (plane_t*)y_hi, (plane_t*)y_lo, H, W, cpad);
  else
    klaunch(nchw_s2d_planes_kernel<SHINEON_FMT_BF16>, grid, 256, 0, (cudaStream_t)stream, x0, C0, x1, C1, (plane_t*)y_hi, (plane_t*)y_lo, H, W, cpad);
  return after_launch("nchw_s2d_planes_kernel");
}

extern "C" int shineon_col2im3x3(const float* t, const float* bias, float* y, int N, int H, int W, int Cout, int tstride,
                                 shineon_stream_t stream) {
  SHINEON_REQUIRE(t && y, "col2im3x3: null pointer");
  SHINEON_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && Cout > 0 && tstride >= 9 * Cout, "col2im3x3: bad shape");
  dim3 grid(grid_x((long)H * W * Cout, 256), N);
  klaunch(col2im3x3_kernel, grid, 256, 0, (cudaStream_t)stream, t, bias, y, H, W, Cout, tstride);
  return after_launch("col2im3x3_kernel");
}

int shineon_upconv3x3_gather_tma(const float* t, const float* bias, float* y, double* stats_ws, int N, int h, int w,
                                 int Cout, int tstride, cudaStream_t stream);  // upconv_gather.cu

static int upconv3x3_gather_direct(const float* t, const float* bias, float* y, int N, int h, int w, int Cout, int tstride,
                                   shineon_stream_t stream);

extern "C" int shineon_upconv3x3_gather(const float* t, const float* bias, float* y, double* stats_ws, int N, int h, int w,
                                        int Cout, int tstride, shineon_stream_t stream) {
  SHINEON_REQUIRE(t && y, "upconv3x3_gather: null pointer");
  SHINEON_REQUIRE(N > 0 && N <= 65535 && h > 0 && w > 0 && Cout > 0 && tstride >= 9 * Cout, "upconv3x3_gather: bad shape");
  {  // TMA-staged variant (upconv_gather.cu) where the channel count allows 32-channel boxes
    const int rc = shineon_upconv3x3_gather_tma(t, bias, y, stats_ws, N, h, w, Cout, tstride, (cudaStream_t)stream);
    if (rc <= 0) return rc;
  }
  int rc = upconv3x3_gather_direct(t, bias, y, N, h, w, Cout, tstride, stream);
  if (rc == SHINEON_OK && stats_ws) rc = launch_instnorm_stats(y, stats_ws, N, 4 * h * w, Cout, (cudaStream_t)stream);
  return rc;
}

static int upconv3x3_gather_direct(const float* t, const float* bias, float* y, int N, int h, int w, int Cout, int tstride,
                                   shineon_stream_t stream) {
  const int vec = (Cout % 4 == 0 && tstride % 4 == 0) ? 4 : 1;
  const int groups = Cout / vec;
  int gpb = 1;  // channel groups per CTA (power of two <= 8); the other 256/gpb threads tile low-res pixels
  while (gpb < 8 && gpb < groups) gpb *= 2;
  const int pixels = 256 / gpb;
  const int tile_w = pixels >= 256 ? 32 : (pixels >= 128 ? 16 : 8);
  const int tile_h = pixels / tile_w;
  const int tiles_x = cdiv(w, tile_w), tiles_y = cdiv(h, tile_h);
  SHINEON_REQUIRE(cdiv(groups, gpb) <= 65535, "upconv3x3_gather: too many channels");
  dim3 grid(tiles_x * tiles_y, cdiv(groups, gpb), N);
  if (vec == 4)
    klaunch(upconv3x3_gather_kernel<4>, grid, 256, 0, (cudaStream_t)stream, t, bias, y, h, w, Cout, tstride, gpb, tile_w, tiles_x);
  else
    klaunch(upconv3x3_gather_kernel<1>, grid, 256, 0, (cudaStream_t)stream, t, bias, y, h, w, Cout, tstride, gpb, tile_w, tiles_x);
  return after_launch("upconv3x3_gather_kernel");
}

// Accumulates per-(n, c) sum / sum of squares of x [N,HW,C] into ws (NOT zeroed here).  Also used by the conv kernel's
// host side for layers whose pixel tiles span several images (conv_igemm.cu).
int launch_instnorm_stats(const float* x, double* ws, int N, int HW, int C, cudaStream_t stream) {
  int cb = 32;
  while (cb > 1 && cb / 2 >= C) cb /= 2;
  const int pp = 256 / cb;
  int pix_per_cta = pp * 16;
  dim3 grid(cdiv(HW, pix_per_cta), cdiv(C, cb), N);
  switch (cb) {
    case 32: klaunch(instnorm_stats_kernel<32>, grid, 256, 0, stream, x, ws, HW, C, pix_per_cta); break;
    case 16: klaunch(instnorm_stats_kernel<16>, grid, 256, 0, stream, x, ws, HW, C, pix_per_cta); break;
    case 8: klaunch(instnorm_stats_kernel<8>, grid, 256, 0, stream, x, ws, HW, C, pix_per_cta); break;
    case 4: klaunch(instnorm_stats_kernel<4>, grid, 256, 0, stream, x, ws, HW, C, pix_per_cta); break;
    case 2: klaunch(instnorm_stats_kernel<2>, grid, 256, 0, stream, x, ws, HW, C, pix_per_cta); break;
    default: klaunch(instnorm_stats_kernel<1>, grid, 256, 0, stream, x, ws, HW, C, pix_per_cta); break;
  }
  return after_launch("instnorm_stats_kernel");
}

extern "C" int shineon_instnorm_act(const float* x, float* y_f32, void* y_hi, void* y_lo, double* stats_ws, int N,
                                    int H, int W, int C, int cpad, float eps, int do_norm, int stats_ready, int act,
                                    float act_param, int plane_fmt, shineon_stream_t stream_) {
  SHINEON_REQUIRE_FMT(plane_fmt, "instnorm_act");
  SHINEON_REQUIRE(x && (y_f32 || y_hi), "instnorm_act: null pointer");
  SHINEON_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && C > 0, "instnorm_act: bad shape");
  SHINEON_REQUIRE(!do_norm || stats_ws, "instnorm_act: stats_ws required");
  SHINEON_REQUIRE(!y_hi || cpad >= C, "instnorm_act: cpad < C");
  SHINEON_REQUIRE(2 * C * sizeof(float) <= 48 * 1024, "instnorm_act: C too large");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int HW = H * W;
  if (do_norm && !stats_ready) {
    cudaError_t e = cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * (size_t)N * C, stream);
    if (e != cudaSuccess) return fail(SHINEON_ERR_CUDA, "instnorm_act memset: %s", cudaGetErrorString(e));
    int rc = launch_instnorm_stats(x, stats_ws, N, HW, C, stream);
    if (rc) return rc;
  }
  // 128-bit plane stores need 16-byte aligned rows: cpad % 8 and 16-byte aligned plane pointers (channel windows are)
  const bool al8 = (C % 8 == 0) && (!y_hi || (cpad % 8 == 0 && ((reinterpret_cast<uintptr_t>(y_hi) | reinterpret_cast<uintptr_t>(y_lo)) & 15) == 0));
  const int vecC = al8 ? 8 : ((C % 4 == 0) ? 4 : 1);
  SHINEON_REQUIRE((long)HW * (C / vecC) < (1l << 31), "instnorm_act: image too large");
  dim3 grid(grid_x((long)HW * (C / vecC), 256), N);
  const size_t sm = 2 * C * sizeof(float);
  cudaStream_t st = stream;
#define SHINEON_IN_APPLY(F, A, V)                                                                                   \
  klaunch(instnorm_apply_kernel<F, A, V>, grid, 256, sm, st, x, stats_ws, y_f32, (plane_t*)y_hi, (plane_t*)y_lo, HW, C, \
                                                         cpad, eps, do_norm, act, act_param)
#define SHINEON_IN_ACT(F, V)                                              \
  switch (act) {                                                          \
    case SHINEON_ACT_NONE: SHINEON_IN_APPLY(F, SHINEON_ACT_NONE, V); break; \
    case SHINEON_ACT_RELU: SHINEON_IN_APPLY(F, SHINEON_ACT_RELU, V); break; \
    case SHINEON_ACT_GELU: SHINEON_IN_APPLY(F, SHINEON_ACT_GELU, V); break; \
    default: SHINEON_IN_APPLY(F, -1, V); break;                           \
  }
  if (plane_fmt == SHINEON_FMT_FP16) {
    if (vecC == 8) { SHINEON_IN_ACT(SHINEON_FMT_FP16, 8) } else if (vecC == 4) { SHINEON_IN_ACT(SHINEON_FMT_FP16, 4) } else { SHINEON_IN_ACT(SHINEON_FMT_FP16, 1) }
  } else {
    if (vecC == 8) { SHINEON_IN_ACT(SHINEON_FMT_BF16, 8) } else if (vecC == 4) { SHINEON_IN_ACT(SHINEON_FMT_BF16, 4) } else { SHINEON_IN_ACT(SHINEON_FMT_BF16, 1) }
  }
#undef SHINEON_IN_ACT
#undef SHINEON_IN_APPLY
  return after_launch("instnorm_apply_kernel");
}

extern "C" int shineon_upsample2x_cat(const void* s0_hi, const void* s0_lo, int c0pad, const void* s1_hi,
                                      const void* s1_lo, int c1pad, void* y_hi, void* y_lo, int N, int H, int W,
                                      int act, float act_param, int plane_fmt, shineon_stream_t stream) {
  SHINEON_REQUIRE_FMT(plane_fmt, "upsample2x_cat");
  SHINEON_REQUIRE(s0_hi && y_hi, "upsample2x_cat: null pointer");
  SHINEON_REQUIRE((s1_hi == nullptr) == (c1pad == 0), "upsample2x_cat: s1/c1pad mismatch");
  SHINEON_REQUIRE((s0_lo == nullptr) == (y_lo == nullptr), "upsample2x_cat: lo planes must be all present or all absent");
  SHINEON_REQUIRE(s1_hi == nullptr || (s1_lo == nullptr) == (s0_lo == nullptr), "upsample2x_cat: s1 lo mismatch");
  SHINEON_REQUIRE(c0pad > 0 && c0pad % 16 == 0 && c1pad % 16 == 0, "upsample2x_cat: channel pads must be multiples of 16");
  SHINEON_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0, "upsample2x_cat: bad shape");
  SHINEON_REQUIRE(H + 1 <= 65535, "upsample2x_cat: H too large");
  dim3 grid(cdiv((W + 1) * ((c0pad + c1pad) / 8), 256), H + 1, N);
#define SHINEON_UP(F, A)                                                                                              \
  klaunch(upsample2x_cat_kernel<F, A>, grid, 256, 0, (cudaStream_t)stream,                                                \
      (const plane_t*)s0_hi, (const plane_t*)s0_lo, c0pad, (const plane_t*)s1_hi, (const plane_t*)s1_lo, c1pad,      \
      (plane_t*)y_hi, (plane_t*)y_lo, H, W, act, act_param)
  if (plane_fmt == SHINEON_FMT_FP16) {
    if (act == SHINEON_ACT_NONE) SHINEON_UP(SHINEON_FMT_FP16, SHINEON_ACT_NONE);
    else if (act == SHINEON_ACT_RELU) SHINEON_UP(SHINEON_FMT_FP16, SHINEON_ACT_RELU);
    else SHINEON_UP(SHINEON_FMT_FP16, -1);
  } else {
    if (act == SHINEON_ACT_NONE) SHINEON_UP(SHINEON_FMT_BF16, SHINEON_ACT_NONE);
    else if (act == SHINEON_ACT_RELU) SHINEON_UP(SHINEON_FMT_BF16, SHINEON_ACT_RELU);
    else SHINEON_UP(SHINEON_FMT_BF16, -1);
  }
#undef SHINEON_UP
  return after_launch("upsample2x_cat_kernel");
}

extern "C" int shineon_tom_compose(const float* unet_out, int Cout, const float* cloth, const float* warped_prev,
                                   float* p_rendereds, float* tryon_masks, float* p_tryons, float* flow_masks,
                                   unsigned char* p_tryons_u8, int B, int H, int W, int n_frames, int frame, int flow_warp,
                                   shineon_stream_t stream) {
  SHINEON_REQUIRE(unet_out && cloth && (p_tryons || p_tryons_u8), "tom_compose: null pointer");
  SHINEON_REQUIRE(n_frames >= 1 && frame >= 0 && frame < n_frames, "tom_compose: frame %d of %d", frame, n_frames);
  SHINEON_REQUIRE(Cout == (flow_warp ? 5 : 4) * n_frames, "tom_compose: Cout %d != %d*n_frames", Cout, flow_warp ? 5 : 4);
  SHINEON_REQUIRE(!warped_prev || flow_warp, "tom_compose: warped_prev needs flow_warp");
  SHINEON_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, "tom_compose: bad shape");
  const int HW = H * W;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec4 = HW % 4 == 0 && al16(cloth) && al16(warped_prev) && al16(p_rendereds) && al16(tryon_masks) &&
                    al16(p_tryons) && al16(flow_masks) && (reinterpret_cast<uintptr_t>(p_tryons_u8) & 3) == 0;
  if (vec4) {
    dim3 grid(grid_x((long)HW / 4, 256), B);
    klaunch(tom_compose_kernel<true>, grid, 256, 0, (cudaStream_t)stream, unet_out, Cout, cloth, warped_prev, p_rendereds, tryon_masks,
                                                                   p_tryons, flow_masks, p_tryons_u8, HW, n_frames, frame, flow_warp);
  } else {
    dim3 grid(grid_x((long)HW, 256), B);
    klaunch(tom_compose_kernel<false>, grid, 256, 0, (cudaStream_t)stream, unet_out, Cout, cloth, warped_prev, p_rendereds, tryon_masks,
                                                                    p_tryons, flow_masks, p_tryons_u8, HW, n_frames, frame, flow_warp);
  }
  return after_launch("tom_compose_kernel");
}

extern "C" int shineon_image_to_u8(const float* x, unsigned char* y, int B, int C, int H, int W, shineon_stream_t stream) {
  SHINEON_REQUIRE(x && y, "image_to_u8: null pointer");
  SHINEON_REQUIRE(B > 0 && B <= 65535 && C > 0 && H > 0 && W > 0, "image_to_u8: bad shape");
  const int HW = H * W;
  const bool vec = HW % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 3) == 0;
  if (vec && C == 3)
    klaunch(image_to_u8_kernel<3>, dim3(cdiv(HW / 4, 256), B), 256, 0, (cudaStream_t)stream, x, y, HW);
  else if (vec && C == 1)
    klaunch(image_to_u8_kernel<1>, dim3(cdiv(HW / 4, 256), B), 256, 0, (cudaStream_t)stream, x, y, HW);
  else
    klaunch(image_to_u8_scalar_kernel, dim3(grid_x((long)HW * C, 256), B), 256, 0, (cudaStream_t)stream, x, y, HW, C);
  return after_launch("image_to_u8_kernel");
}
