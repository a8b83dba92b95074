// FlowNet cost-volume correlation (reference: correlation_cuda_kernel.cu:74-147 forward,
// :151-334 backward; shape rules correlation_cuda.cc:19-38).
//
// out[n, tj*D+ti, oy, ox] = 1/(k*k*C) * sum_{j,i in window} sum_c
//        pad(in1)[n, c, y1+j, x1+i] * pad(in2)[n, c, y1+tj*s2+j, x1+ti*s2+i]
// with (y1,x1) = (oy,ox)*s1 + max_displacement in zero-padded coordinates, D = 2*(maxd/s2)+1.
//
// Unlike the reference this works straight from the NCHW inputs: no padded NHWC copies
// (rbot1/rbot2, correlation_cuda.cc:36-42) are materialised; out-of-image taps contribute zero.
#include "common.cuh"

namespace shineon {

struct CorrGeom {
  int C, H, W, pad, k, maxd, s1, s2;
  int kr, drad, D, outC, outH, outW;
};

static inline bool corr_geom(int C, int H, int W, int pad, int k, int maxd, int s1, int s2, CorrGeom& g) {
  if (C <= 0 || H <= 0 || W <= 0 || pad < 0 || k <= 0 || (k & 1) == 0 || maxd < 0 || s1 <= 0 || s2 <= 0) return false;
  g.C = C; g.H = H; g.W = W; g.pad = pad; g.k = k; g.maxd = maxd; g.s1 = s1; g.s2 = s2;
  g.kr = (k - 1) / 2;
  int border = g.kr + maxd;
  int pH = H + 2 * pad, pW = W + 2 * pad;
  g.drad = maxd / s2;
  g.D = 2 * g.drad + 1;
  g.outC = g.D * g.D;
  // ceil(float(pH - 2*border)/float(s1))  (correlation_cuda.cc:33-34)
  g.outH = (int)ceilf((float)(pH - 2 * border) / (float)s1);
  g.outW = (int)ceilf((float)(pW - 2 * border) / (float)s1);
  return g.outH > 0 && g.outW > 0;
}

// One CTA per (n, oy, block of displacement rows); threads: x fastest so global reads coalesce along W.
__global__ void __launch_bounds__(256)
    correlation_fwd_kernel(const float* __restrict__ in1, const float* __restrict__ in2, float* __restrict__ out,
                           CorrGeom g) {
  pdl_grid_sync();
  const int n = blockIdx.z, oy = blockIdx.y;
  const long HW = (long)g.H * g.W;
  const float* a = in1 + (long)n * g.C * HW;
  const float* b = in2 + (long)n * g.C * HW;
  const int per_row = g.D * g.outW;  // (ti, ox) pairs for one tj
  const float inv = 1.f / (float)(g.k * g.k * g.C);
  for (int tj = blockIdx.x; tj < g.D; tj += gridDim.x) {
    for (int idx = threadIdx.x; idx < per_row; idx += blockDim.x) {
      const int ti = idx / g.outW, ox = idx - ti * g.outW;
      // image (unpadded) coordinates of the window centres
      const int y1 = oy * g.s1 + g.maxd - g.pad, x1 = ox * g.s1 + g.maxd - g.pad;
      const int y2 = y1 + (tj - g.drad) * g.s2, x2 = x1 + (ti - g.drad) * g.s2;
      float acc = 0.f;
      for (int j = -g.kr; j <= g.kr; ++j) {
        const int ya = y1 + j, yb = y2 + j;
        if (ya < 0 || ya >= g.H || yb < 0 || yb >= g.H) continue;
        for (int i = -g.kr; i <= g.kr; ++i) {
          const int xa = x1 + i, xb = x2 + i;
          if (xa < 0 || xa >= g.W || xb < 0 || xb >= g.W) continue;
          const float* pa = a + (long)ya * g.W + xa;
          const float* pb = b + (long)yb * g.W + xb;
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          int c = 0;
          for (; c + 3 < g.C; c += 4) {
            s0 = fmaf(__ldg(pa + (c + 0) * HW), __ldg(pb + (c + 0) * HW), s0);
            s1 = fmaf(__ldg(pa + (c + 1) * HW), __ldg(pb + (c + 1) * HW), s1);
            s2 = fmaf(__ldg(pa + (c + 2) * HW), __ldg(pb + (c + 2) * HW), s2);
            s3 = fmaf(__ldg(pa + (c + 3) * HW), __ldg(pb + (c + 3) * HW), s3);
          }
          for (; c < g.C; ++c) s0 = fmaf(__ldg(pa + c * HW), __ldg(pb + c * HW), s0);
          acc += (s0 + s1) + (s2 + s3);
        }
      }
      const int tc = tj * g.D + ti;
      out[(((long)n * g.outC + tc) * g.outH + oy) * g.outW + ox] = acc * inv;
    }
  }
}

// Gather of the tensor-core path: full[b, p1, p2] = <in1[p1,:], in2[p2,:]> (all pairs, one GEMM per image) ->
// out[b, tj*D+ti, y, x] = full[b, (y,x), (y + (tj-r)*s2, x + (ti-r)*s2)] / C, zero outside the image.
__global__ void __launch_bounds__(256)
    correlation_gather_kernel(const float* __restrict__ full, float* __restrict__ out, int C, int H, int W, int shift,
                              int drad, int s2, long total) {
  pdl_grid_sync();
  const int D = 2 * drad + 1, P = H * W;
  const float inv = 1.f / (float)C;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int x = (int)(e % W), y = (int)((e / W) % H);
    const int tc = (int)((e / P) % (D * D)), b = (int)(e / ((long)P * D * D));
    const int tj = tc / D, ti = tc - tj * D;
    const int y1 = y + shift, x1 = x + shift;  // window centre in image coordinates (shift = maxd - pad)
    const int y2 = y1 + (tj - drad) * s2, x2 = x1 + (ti - drad) * s2;
    float v = 0.f;
    if (y1 >= 0 && y1 < H && x1 >= 0 && x1 < W && y2 >= 0 && y2 < H && x2 >= 0 && x2 < W)
      v = __ldg(full + ((long)b * P + (y1 * W + x1)) * P + (y2 * W + x2)) * inv;
    out[e] = v;
  }
}

// The same gather for a consumer that wants NHWC 16-bit planes (FlowNetC: LeakyReLU(cost volume) is a channel window of
// conv3_1's concat input).  One warp per output pixel: the pixel's row of the all-pairs matrix (H*W floats, L2-resident
// right after the GEMM) is read by the whole warp, and the D*D displacement values leave as contiguous 16-byte chunks of
// the pixel's channel vector with the activation and the hi/lo split applied -- instead of an NCHW f32 cost volume
// (element stride H*W between a pixel's displacements) that a second kernel re-reads, transposes and splits.
template <int FMT>
__global__ void __launch_bounds__(256)
    correlation_gather_planes_kernel(const float* __restrict__ full, plane_t* __restrict__ yh, plane_t* __restrict__ yl,
                                     int C, int H, int W, int shift, int drad, int s2, int cstride, int cpad, int act,
                                     float act_param, long pixels) {
  pdl_grid_sync();
  const int D = 2 * drad + 1, DD = D * D, P = H * W;
  const float inv = 1.f / (float)C;
  const int lane = threadIdx.x & 31;
  for (long wp = (long)blockIdx.x * 8 + (threadIdx.x >> 5); wp < pixels; wp += (long)gridDim.x * 8) {
    const int p = (int)(wp % P);
    const int y = p / W, x = p - y * W;
    const int y1 = y + shift, x1 = x + shift;
    const bool ok1 = y1 >= 0 && y1 < H && x1 >= 0 && x1 < W;
    const float* row = full + ((wp - p) + (long)(y1 * W + x1)) * P;  // row of pixel (b, y1, x1)
    for (int ch = lane; ch < (cpad >> 3); ch += 32) {
      float v[8];
      int d = ch * 8;
      int tj = d / D, ti = d - tj * D;
#pragma unroll
      for (int j = 0; j < 8; ++j, ++d) {
        float t = 0.f;
        if (d < DD) {
          const int y2 = y1 + (tj - drad) * s2, x2 = x1 + (ti - drad) * s2;
          if (ok1 && y2 >= 0 && y2 < H && x2 >= 0 && x2 < W) t = __ldg(row + (y2 * W + x2)) * inv;
          t = apply_act(t, act, act_param);
        }
        v[j] = t;
        if (++ti == D) { ti = 0; ++tj; }
      }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split16x2(v[2 * j], v[2 * j + 1], FMT, hi[j], lo[j]);
      const long o = wp * cstride + ch * 8;
      *reinterpret_cast<uint4*>(yh + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      if (yl) *reinterpret_cast<uint4*>(yl + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// Backward, one thread per input element, loops follow correlation_cuda_kernel.cu:151-241 / :244-334
// (integer divisions truncate toward zero exactly like the reference).
__global__ void __launch_bounds__(256)
    correlation_bwd_kernel(const float* __restrict__ in1, const float* __restrict__ in2,
                           const float* __restrict__ gout, float* __restrict__ gin1, float* __restrict__ gin2,
                           CorrGeom g, long total) {
  pdl_grid_sync();
  const long HW = (long)g.H * g.W;
  const float nelems = (float)(g.k * g.k * g.C);
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int xi = (int)(e % g.W);
    const int yi = (int)((e / g.W) % g.H);
    const int c = (int)((e / HW) % g.C);
    const int n = (int)(e / (HW * g.C));
    const int y = yi + g.pad, x = xi + g.pad;  // padded coordinates
    const float* go = gout + (long)n * g.outC * g.outH * g.outW;
    // ---- grad wrt input1 (:170-240)
    float acc1 = 0.f;
    {
      int xmin = (x - g.kr - g.maxd) / g.s1, ymin = (y - g.kr - g.maxd) / g.s1;
      int xmax = (x + g.kr - g.maxd) / g.s1, ymax = (y + g.kr - g.maxd) / g.s1;
      bool skip = (xmax < 0 || ymax < 0 || xmin >= g.outW || ymin >= g.outH) || (xmin > xmax || ymin > ymax);
      if (!skip) {
        xmin = max(0, xmin); xmax = min(g.outW - 1, xmax);
        ymin = max(0, ymin); ymax = min(g.outH - 1, ymax);
        for (int tc = 0; tc < g.outC; ++tc) {
          int i2 = (tc % g.D - g.drad) * g.s2, j2 = (tc / g.D - g.drad) * g.s2;
          int yy = y + j2 - g.pad, xx = x + i2 - g.pad;
          float v2 = (yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) ? in2[((long)n * g.C + c) * HW + (long)yy * g.W + xx] : 0.f;
          if (v2 == 0.f) continue;
          for (int j = ymin; j <= ymax; ++j)
            for (int i = xmin; i <= xmax; ++i) acc1 += go[((long)tc * g.outH + j) * g.outW + i] * v2;
        }
      }
    }
    gin1[e] = acc1 / nelems;
    // ---- grad wrt input2 (:263-333)
    float acc2 = 0.f;
    for (int tc = 0; tc < g.outC; ++tc) {
      int i2 = (tc % g.D - g.drad) * g.s2, j2 = (tc / g.D - g.drad) * g.s2;
      int xmin = (x - g.kr - g.maxd - i2) / g.s1, ymin = (y - g.kr - g.maxd - j2) / g.s1;
      int xmax = (x + g.kr - g.maxd - i2) / g.s1, ymax = (y + g.kr - g.maxd - j2) / g.s1;
      if (xmax < 0 || ymax < 0 || xmin >= g.outW || ymin >= g.outH) continue;
      if (xmin > xmax || ymin > ymax) continue;
      xmin = max(0, xmin); xmax = min(g.outW - 1, xmax);
      ymin = max(0, ymin); ymax = min(g.outH - 1, ymax);
      int yy = y - j2 - g.pad, xx = x - i2 - g.pad;
      float v1 = (yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) ? in1[((long)n * g.C + c) * HW + (long)yy * g.W + xx] : 0.f;
      if (v1 == 0.f) continue;
      for (int j = ymin; j <= ymax; ++j)
        for (int i = xmin; i <= xmax; ++i) acc2 += go[((long)tc * g.outH + j) * g.outW + i] * v1;
    }
    gin2[e] = acc2 / nelems;
  }
}

}  // namespace shineon

using namespace shineon;

extern "C" int shineon_correlation_out_shape(int C, int H, int W, int pad_size, int kernel_size, int max_displacement,
                                             int stride1, int stride2, int* out_c, int* out_h, int* out_w) {
  CorrGeom g;
  if (!corr_geom(C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2, g))
    return fail(SHINEON_ERR_ARG, "correlation: bad geometry C=%d H=%d W=%d pad=%d k=%d maxd=%d s1=%d s2=%d", C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2);
  if (out_c) *out_c = g.outC;
  if (out_h) *out_h = g.outH;
  if (out_w) *out_w = g.outW;
  return SHINEON_OK;
}

extern "C" int shineon_correlation_fwd(const float* in1, const float* in2, float* out, int B, int C, int H, int W,
                                       int pad_size, int kernel_size, int max_displacement, int stride1, int stride2,
                                       shineon_stream_t stream) {
  SHINEON_REQUIRE(in1 && in2 && out, "correlation_fwd: null pointer");
  CorrGeom g;
  if (!corr_geom(C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2, g))
    return fail(SHINEON_ERR_ARG, "correlation_fwd: bad geometry");
  SHINEON_REQUIRE(B >= 0 && B <= 65535 && g.outH <= 65535, "correlation_fwd: bad batch");
  if (B == 0) return SHINEON_OK;
  dim3 grid(g.D, g.outH, B);
  klaunch(correlation_fwd_kernel, grid, 256, 0, (cudaStream_t)stream, in1, in2, out, g);
  return after_launch("correlation_fwd_kernel");
}

extern "C" int shineon_correlation_gather(const float* full, float* out, int B, int C, int H, int W, int pad_size,
                                          int max_displacement, int stride2, shineon_stream_t stream) {
  SHINEON_REQUIRE(full && out, "correlation_gather: null pointer");
  CorrGeom g;
  if (!corr_geom(C, H, W, pad_size, 1, max_displacement, 1, stride2, g))
    return fail(SHINEON_ERR_ARG, "correlation_gather: bad geometry");
  SHINEON_REQUIRE(g.outH == H && g.outW == W, "correlation_gather: needs pad_size == max_displacement (output size == input size)");
  const long total = (long)B * g.outC * H * W;
  if (total == 0) return SHINEON_OK;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  klaunch(correlation_gather_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, full, out, C, H, W, max_displacement - pad_size,
                                                                          g.drad, stride2, total);
  return after_launch("correlation_gather_kernel");
}

extern "C" int shineon_correlation_gather_planes(const float* full, void* y_hi, void* y_lo, int y_cstride, int B, int C, int H,
                                                 int W, int pad_size, int max_displacement, int stride2, int act, float act_param,
                                                 int plane_fmt, shineon_stream_t stream) {
  SHINEON_REQUIRE(full && y_hi, "correlation_gather_planes: null pointer");
  SHINEON_REQUIRE(plane_fmt == SHINEON_FMT_BF16 || plane_fmt == SHINEON_FMT_FP16, "correlation_gather_planes: plane_fmt %d", plane_fmt);
  CorrGeom g;
  if (!corr_geom(C, H, W, pad_size, 1, max_displacement, 1, stride2, g))
    return fail(SHINEON_ERR_ARG, "correlation_gather_planes: bad geometry");
  SHINEON_REQUIRE(g.outH == H && g.outW == W, "correlation_gather_planes: needs pad_size == max_displacement (output size == input size)");
  const int cpad = (g.outC + 7) / 8 * 8;
  SHINEON_REQUIRE(y_cstride >= cpad && y_cstride % 8 == 0, "correlation_gather_planes: y_cstride %d too small for %d channels / not a multiple of 8", y_cstride, g.outC);
  SHINEON_REQUIRE((reinterpret_cast<uintptr_t>(y_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(y_lo) & 15) == 0, "correlation_gather_planes: planes must be 16-byte aligned");
  const long pixels = (long)B * H * W;
  if (pixels == 0) return SHINEON_OK;
  long blocks = (pixels + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (plane_fmt == SHINEON_FMT_FP16)
    klaunch(correlation_gather_planes_kernel<SHINEON_FMT_FP16>, (int)blocks, 256, 0, (cudaStream_t)stream, full, (plane_t*)y_hi,
            (plane_t*)y_lo, C, H, W, max_displacement - pad_size, g.drad, stride2, y_cstride, cpad, act, act_param, pixels);
  else
    klaunch(correlation_gather_planes_kernel<SHINEON_FMT_BF16>, (int)blocks, 256, 0, (cudaStream_t)stream, full, (plane_t*)y_hi,
            (plane_t*)y_lo, C, H, W, max_displacement - pad_size, g.drad, stride2, y_cstride, cpad, act, act_param, pixels);
  return after_launch("correlation_gather_planes_kernel");
}

extern "C" int shineon_correlation_bwd(const float* in1, const float* in2, const float* grad_out, float* grad_in1,
                                       float* grad_in2, int B, int C, int H, int W, int pad_size, int kernel_size,
                                       int max_displacement, int stride1, int stride2, shineon_stream_t stream) {
  SHINEON_REQUIRE(in1 && in2 && grad_out && grad_in1 && grad_in2, "correlation_bwd: null pointer");
  CorrGeom g;
  if (!corr_geom(C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2, g))
    return fail(SHINEON_ERR_ARG, "correlation_bwd: bad geometry");
  long total = (long)B * C * H * W;
  if (total == 0) return SHINEON_OK;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  klaunch(correlation_bwd_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, in1, in2, grad_out, grad_in1, grad_in2, g, total);
  return after_launch("correlation_bwd_kernel");
}
