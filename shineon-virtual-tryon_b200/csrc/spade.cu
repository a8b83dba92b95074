// Memory-bound passes of the SAMS generator, all NHWC:
//   * spade_modulate        : SPADE.forward's `normalized * (1 + gamma) + beta` (models/networks/sams/spade.py:68-84) fused
//                             with the parameter-free norm (instance statistics or eval-mode batch-norm affine), the
//                             activation AnySpadeResBlock applies next (spade.py:157-158) and the hi/lo split of the conv operand
//   * nearest_resize_nhwc   : nn.Upsample(scale_factor=0.5 / 2) between the blocks (sams_generator.py:295-310)
//   * nearest_resize_planes : F.interpolate(segmap, size, mode="nearest") (spade.py:74) written as the 16-bit conv operand
//   * add_nhwc              : the residual `x_s + dx` (spade.py:160)
//   * chan_stats            : per-(image, channel) sum / sum of squares for the instance-norm flavour
#include "common.cuh"

int launch_instnorm_stats(const float* x, double* ws, int N, int HW, int C, cudaStream_t stream);  // norm_act.cu

namespace shineon {

static inline int grid_1d(long total, int threads) {
  long g = (total + threads - 1) / threads;
  const long cap = 148l * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// PyTorch's nearest_neighbor_compute_source_index (UpSample.h): min(floor(dst * scale), in - 1)
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
  return min((int)floorf((float)dst * scale), in_size - 1);
}

// norm_mode: 0 none, 1 instance statistics (ws = f64 [N][C][2] sum / sum of squares over HW), 2 per-channel affine
// (nscale / nshift f32 [C]: eval-mode BatchNorm2d(affine=False) = 1/sqrt(var+eps), -mean/sqrt(var+eps)).
// gb: f32 NHWC [N,H,W,gb_cstride] holding (1 + gamma) in channels [0,C) and beta in [C,2C) (one conv, bias + 1 folded in).
template <int FMT, int VEC>
__global__ void __launch_bounds__(256)
    spade_modulate_kernel(const float* __restrict__ x, const double* __restrict__ ws, const float* __restrict__ nscale,
                          const float* __restrict__ nshift, const float* __restrict__ gb, int gb_cstride,
                          float* __restrict__ yf, int yf_cstride, plane_t* __restrict__ yh, plane_t* __restrict__ yl,
                          int HW, int C, int cpad, float eps, int norm_mode, int act, float act_param) {
  pdl_grid_sync();
  extern __shared__ float s_tab[];  // a[C], b[C]: normalized = x * a + b
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 1.f, b = 0.f;
    if (norm_mode == 1) {
      const double s = ws[((long)n * C + c) * 2], q = ws[((long)n * C + c) * 2 + 1];
      const double m = s / HW;
      double var = q / HW - m * m;  // biased variance (F.instance_norm)
      if (var < 0.0) var = 0.0;
      const double r = 1.0 / sqrt(var + (double)eps);
      a = (float)r;
      b = (float)(-m * r);
    } else if (norm_mode == 2) {
      a = nscale[c];
      b = nshift[c];
    }
    s_tab[c] = a;
    s_tab[C + c] = b;
  }
  __syncthreads();
  const int cg = C / VEC;
  const unsigned total = (unsigned)HW * (unsigned)cg;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const unsigned p = e / (unsigned)cg;
    const int c0 = (int)(e - p * (unsigned)cg) * VEC;
    const long pix = (long)n * HW + p;
    float v[VEC], g[VEC], bt[VEC];
    if constexpr (VEC == 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(x + pix * C + c0));
      const float4 tg = __ldg(reinterpret_cast<const float4*>(gb + pix * gb_cstride + c0));
      const float4 tb = __ldg(reinterpret_cast<const float4*>(gb + pix * gb_cstride + C + c0));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      g[0] = tg.x; g[1] = tg.y; g[2] = tg.z; g[3] = tg.w;
      bt[0] = tb.x; bt[1] = tb.y; bt[2] = tb.z; bt[3] = tb.w;
    } else {
      v[0] = x[pix * C + c0];
      g[0] = gb[pix * gb_cstride + c0];
      bt[0] = gb[pix * gb_cstride + C + c0];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      // the reference's operation order: normalise (x - mean) * rstd, then * (1 + gamma) + beta, each rounded to f32
      const float nrm = fmaf(v[j], s_tab[c0 + j], s_tab[C + c0 + j]);
      v[j] = apply_act(nrm * g[j] + bt[j], act, act_param);
    }
    if (yf) {
      if constexpr (VEC == 4)
        *reinterpret_cast<float4*>(yf + pix * yf_cstride + c0) = make_float4(v[0], v[1], v[2], v[3]);
      else
        yf[pix * yf_cstride + c0] = v[0];
    }
    if (yh) {
      const long po = pix * cpad + c0;
      if constexpr (VEC == 4) {
        uint32_t h01, h23, l01, l23;
        split16x2(v[0], v[1], FMT, h01, l01);
        split16x2(v[2], v[3], FMT, h23, l23);
        *reinterpret_cast<uint2*>(yh + po) = make_uint2(h01, h23);
        if (yl) *reinterpret_cast<uint2*>(yl + po) = make_uint2(l01, l23);
      } else {
        plane_t h, l;
        split16(v[0], FMT, h, l);
        yh[po] = h;
        if (yl) yl[po] = l;
      }
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256)
    nearest_resize_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int Hs, int Ws, int H, int W, int C,
                               float scale_h, float scale_w, long total) {
  pdl_grid_sync();
  const int cg = C / VEC;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int g = (int)(e % cg);
    long r = e / cg;
    const int w = (int)(r % W);
    r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    const int sh = nearest_src(h, scale_h, Hs), sw = nearest_src(w, scale_w, Ws);
    const float* src = x + (((long)n * Hs + sh) * Ws + sw) * C + g * VEC;
    float* dst = y + e * VEC;
    if constexpr (VEC == 4)
      *reinterpret_cast<float4*>(dst) = __ldg(reinterpret_cast<const float4*>(src));
    else
      dst[0] = src[0];
  }
}

__global__ void __launch_bounds__(256)
    nearest_resize_planes_kernel(const float* __restrict__ x, int C, int Hs, int Ws, plane_t* __restrict__ yh,
                                 plane_t* __restrict__ yl, int H, int W, int cpad, float scale_h, float scale_w, int fmt) {
  pdl_grid_sync();
  const int n = blockIdx.y;
  const int groups = cpad >> 3;
  const long total = (long)H * W * groups;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int p = (int)(e % ((long)H * W));  // pixel fastest: neighbouring threads read neighbouring source columns
    const int g = (int)(e / ((long)H * W));
    const int h = p / W, w = p - h * W;
    const int sh = nearest_src(h, scale_h, Hs), sw = nearest_src(w, scale_w, Ws);
    __align__(16) plane_t hi[8];
    __align__(16) plane_t lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      const float v = c < C ? __ldg(x + (((long)n * C + c) * Hs + sh) * Ws + sw) : 0.f;
      split16(v, fmt, hi[j], lo[j]);
    }
    const long o = ((long)n * H * W + p) * cpad + g * 8;
    *reinterpret_cast<uint4*>(yh + o) = *reinterpret_cast<const uint4*>(hi);
    if (yl) *reinterpret_cast<uint4*>(yl + o) = *reinterpret_cast<const uint4*>(lo);
  }
}

// im2col (ks x ks, stride 1, pad ks/2) of the nearest-resized label map, straight from the source map: the operand of the
// mlp_shared conv as a dense 1x1 GEMM with K = pad64(ks*ks*C) (label maps have 2-18 channels: one 64-channel K-block per
// filter tap would be 9x the MMA work).  k = (fy*ks + fx)*C + c like nchw_im2col_planes_kernel.
__global__ void __launch_bounds__(256)
    nearest_im2col_planes_kernel(const float* __restrict__ x, int C, int Hs, int Ws, plane_t* __restrict__ yh,
                                 plane_t* __restrict__ yl, int H, int W, int ks, int kpad, float scale_h, float scale_w, int fmt) {
  pdl_grid_sync();
  const int n = blockIdx.y;
  const int groups = kpad >> 3;
  const int pad = ks / 2, K = ks * ks * C;
  const long total = (long)H * W * groups;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int p = (int)(e % ((long)H * W));
    const int g = (int)(e / ((long)H * W));
    const int h = p / W, w = p - h * W;
    __align__(16) plane_t hi[8];
    __align__(16) plane_t lo[8];
    int k = g * 8;
    int tap = k / C, c = k - tap * C;
    int fy = tap / ks, fx = tap - fy * ks;
#pragma unroll
    for (int j = 0; j < 8; ++j, ++k) {
      float v = 0.f;
      const int hh = h + fy - pad, ww = w + fx - pad;  // position in the RESIZED map (zero padding outside it)
      if (k < K && hh >= 0 && hh < H && ww >= 0 && ww < W)
        v = __ldg(x + (((long)n * C + c) * Hs + nearest_src(hh, scale_h, Hs)) * Ws + nearest_src(ww, scale_w, Ws));
      split16(v, fmt, hi[j], lo[j]);
      if (++c == C) {
        c = 0;
        if (++fx == ks) { fx = 0; ++fy; }
      }
    }
    const long o = ((long)n * H * W + p) * kpad + g * 8;
    *reinterpret_cast<uint4*>(yh + o) = *reinterpret_cast<const uint4*>(hi);
    if (yl) *reinterpret_cast<uint4*>(yl + o) = *reinterpret_cast<const uint4*>(lo);
  }
}

// y may alias a or b (every element is read and written by the same thread): no __restrict__, no read-only loads
__global__ void __launch_bounds__(256) add_nhwc_kernel(const float* a, const float* b, float* y, long n4, long n) {
  pdl_grid_sync();
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long)gridDim.x * blockDim.x) {
    const float4 u = reinterpret_cast<const float4*>(a)[e], v = reinterpret_cast<const float4*>(b)[e];
    reinterpret_cast<float4*>(y)[e] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
  }
  if (blockIdx.x == 0)
    for (long e = n4 * 4 + threadIdx.x; e < n; e += blockDim.x) y[e] = a[e] + b[e];
}

// SamsModel.generate_n_frames' tail (models/sams_model.py:226-236): split the generator output into the frame and the
// blend weight, fake = (1 - w) * warped_prev + w * frame (flow_warp) or the frame itself; NHWC in, NCHW frame slot out.
__global__ void __launch_bounds__(256)
    sams_flow_blend_kernel(const float* __restrict__ g, int Cg, const float* __restrict__ warped, float* __restrict__ out,
                           long out_bstride, int HW) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= HW) return;
  const float* gp = g + ((long)b * HW + p) * Cg;
  float r[3] = {gp[0], gp[1], gp[2]};
  if (warped) {
    const float w = gp[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) r[c] = (1.f - w) * __ldg(warped + ((long)b * 3 + c) * HW + p) + w * r[c];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) out[(long)b * out_bstride + (long)c * HW + p] = r[c];
}

}  // namespace shineon

using namespace shineon;

extern "C" int shineon_sams_flow_blend(const float* gen_out, int Cg, const float* warped_prev, float* out, long out_bstride,
                                       int B, int H, int W, shineon_stream_t stream) {
  SHINEON_REQUIRE(gen_out && out && B > 0 && B <= 65535 && H > 0 && W > 0, "sams_flow_blend: bad argument");
  SHINEON_REQUIRE(Cg == (warped_prev ? 4 : 3), "sams_flow_blend: generator output has %d channels, expected %d", Cg, warped_prev ? 4 : 3);
  SHINEON_REQUIRE(out_bstride >= 3l * H * W, "sams_flow_blend: out_bstride");
  dim3 grid(cdiv(H * W, 256), B);
  klaunch(sams_flow_blend_kernel, grid, 256, 0, (cudaStream_t)stream, gen_out, Cg, warped_prev, out, out_bstride, H * W);
  return after_launch("sams_flow_blend_kernel");
}

extern "C" int shineon_chan_stats(const float* x, double* stats_ws, int N, int HW, int C, shineon_stream_t stream_) {
  SHINEON_REQUIRE(x && stats_ws && N > 0 && N <= 65535 && HW > 0 && C > 0, "chan_stats: bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  cudaError_t e = cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * (size_t)N * C, stream);
  if (e != cudaSuccess) return fail(SHINEON_ERR_CUDA, "chan_stats memset: %s", cudaGetErrorString(e));
  return launch_instnorm_stats(x, stats_ws, N, HW, C, stream);
}

extern "C" int shineon_spade_modulate(const float* x, const double* stats_ws, const float* nscale, const float* nshift,
                                      const float* gb, int gb_cstride, float* y_f32, int y_cstride, void* y_hi, void* y_lo,
                                      int N, int H, int W, int C, int cpad, float eps, int norm_mode, int act,
                                      float act_param, int plane_fmt, shineon_stream_t stream) {
  SHINEON_REQUIRE(plane_fmt == SHINEON_FMT_BF16 || plane_fmt == SHINEON_FMT_FP16, "spade_modulate: plane_fmt %d", plane_fmt);
  SHINEON_REQUIRE(x && gb && (y_f32 || y_hi), "spade_modulate: null pointer");
  SHINEON_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && C > 0, "spade_modulate: bad shape");
  SHINEON_REQUIRE(gb_cstride >= 2 * C, "spade_modulate: gamma|beta tensor has %d channels, need %d", gb_cstride, 2 * C);
  SHINEON_REQUIRE(norm_mode >= 0 && norm_mode <= 2, "spade_modulate: norm_mode %d", norm_mode);
  SHINEON_REQUIRE(norm_mode != 1 || stats_ws, "spade_modulate: instance statistics required");
  SHINEON_REQUIRE(norm_mode != 2 || (nscale && nshift), "spade_modulate: per-channel scale / shift required");
  SHINEON_REQUIRE(!y_hi || cpad >= C, "spade_modulate: cpad < C");
  SHINEON_REQUIRE(!y_f32 || y_cstride >= C, "spade_modulate: y_cstride < C");
  SHINEON_REQUIRE(2 * (size_t)C * sizeof(float) <= 48 * 1024, "spade_modulate: C too large");
  const long HW = (long)H * W;
  const auto al = [](const void* p, int a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
  const bool v4 = C % 4 == 0 && gb_cstride % 4 == 0 && al(x, 16) && al(gb, 16) && (!y_f32 || (y_cstride % 4 == 0 && al(y_f32, 16))) &&
                  (!y_hi || (cpad % 4 == 0 && al(y_hi, 8) && al(y_lo, 8)));
  const int vec = v4 ? 4 : 1;
  SHINEON_REQUIRE(HW * (C / vec) < (1l << 31), "spade_modulate: image too large");
  dim3 grid(grid_1d(HW * (C / vec), 256), N);
  const size_t sm = 2 * (size_t)C * sizeof(float);
#define SHINEON_SPADE(F, V)                                                                                               \
  klaunch(spade_modulate_kernel<F, V>, grid, 256, sm, (cudaStream_t)stream, x, stats_ws, nscale, nshift, gb, gb_cstride, y_f32, \
                                                                       y_cstride, (plane_t*)y_hi, (plane_t*)y_lo, (int)HW, C, \
                                                                       cpad, eps, norm_mode, act, act_param)
  if (plane_fmt == SHINEON_FMT_FP16) {
    if (v4) SHINEON_SPADE(SHINEON_FMT_FP16, 4); else SHINEON_SPADE(SHINEON_FMT_FP16, 1);
  } else {
    if (v4) SHINEON_SPADE(SHINEON_FMT_BF16, 4); else SHINEON_SPADE(SHINEON_FMT_BF16, 1);
  }
#undef SHINEON_SPADE
  return after_launch("spade_modulate_kernel");
}

extern "C" int shineon_nearest_resize_nhwc(const float* x, float* y, int N, int Hs, int Ws, int H, int W, int C,
                                           float scale_h, float scale_w, shineon_stream_t stream) {
  SHINEON_REQUIRE(x && y && N > 0 && Hs > 0 && Ws > 0 && H > 0 && W > 0 && C > 0, "nearest_resize_nhwc: bad argument");
  SHINEON_REQUIRE(scale_h > 0.f && scale_w > 0.f, "nearest_resize_nhwc: bad scale");
  const bool v4 = C % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  const long total = (long)N * H * W * (C / (v4 ? 4 : 1));
  if (v4)
    klaunch(nearest_resize_nhwc_kernel<4>, grid_1d(total, 256), 256, 0, (cudaStream_t)stream, x, y, Hs, Ws, H, W, C, scale_h, scale_w, total);
  else
    klaunch(nearest_resize_nhwc_kernel<1>, grid_1d(total, 256), 256, 0, (cudaStream_t)stream, x, y, Hs, Ws, H, W, C, scale_h, scale_w, total);
  return after_launch("nearest_resize_nhwc_kernel");
}

extern "C" int shineon_nearest_resize_planes(const float* x, int N, int C, int Hs, int Ws, void* y_hi, void* y_lo, int H,
                                             int W, int cpad, float scale_h, float scale_w, int plane_fmt,
                                             shineon_stream_t stream) {
  SHINEON_REQUIRE(plane_fmt == SHINEON_FMT_BF16 || plane_fmt == SHINEON_FMT_FP16, "nearest_resize_planes: plane_fmt %d", plane_fmt);
  SHINEON_REQUIRE(x && y_hi && N > 0 && N <= 65535 && C > 0 && Hs > 0 && Ws > 0 && H > 0 && W > 0, "nearest_resize_planes: bad argument");
  SHINEON_REQUIRE(cpad % 8 == 0 && cpad >= C, "nearest_resize_planes: cpad %d too small / not a multiple of 8", cpad);
  SHINEON_REQUIRE(scale_h > 0.f && scale_w > 0.f, "nearest_resize_planes: bad scale");
  dim3 grid(grid_1d((long)H * W * (cpad >> 3), 256), N);
  klaunch(nearest_resize_planes_kernel, grid, 256, 0, (cudaStream_t)stream, x, C, Hs, Ws, (plane_t*)y_hi, (plane_t*)y_lo, H, W, cpad,
                                                                       scale_h, scale_w, plane_fmt);
  return after_launch("nearest_resize_planes_kernel");
}

extern "C" int shineon_nearest_im2col_planes(const float* x, int N, int C, int Hs, int Ws, void* y_hi, void* y_lo, int H, int W,
                                            int ks, int kpad, float scale_h, float scale_w, int plane_fmt,
                                            shineon_stream_t stream) {
  SHINEON_REQUIRE(plane_fmt == SHINEON_FMT_BF16 || plane_fmt == SHINEON_FMT_FP16, "nearest_im2col_planes: plane_fmt %d", plane_fmt);
  SHINEON_REQUIRE(x && y_hi && N > 0 && N <= 65535 && C > 0 && Hs > 0 && Ws > 0 && H > 0 && W > 0, "nearest_im2col_planes: bad argument");
  SHINEON_REQUIRE(ks >= 1 && ks <= 7 && (ks & 1) == 1, "nearest_im2col_planes: kernel size %d", ks);
  SHINEON_REQUIRE(kpad % 8 == 0 && kpad >= ks * ks * C, "nearest_im2col_planes: kpad %d too small / not a multiple of 8", kpad);
  SHINEON_REQUIRE(scale_h > 0.f && scale_w > 0.f, "nearest_im2col_planes: bad scale");
  dim3 grid(grid_1d((long)H * W * (kpad >> 3), 256), N);
  klaunch(nearest_im2col_planes_kernel, grid, 256, 0, (cudaStream_t)stream, x, C, Hs, Ws, (plane_t*)y_hi, (plane_t*)y_lo, H, W, ks, kpad,
                                                                       scale_h, scale_w, plane_fmt);
  return after_launch("nearest_im2col_planes_kernel");
}

extern "C" int shineon_add_nhwc(const float* a, const float* b, float* y, long n, shineon_stream_t stream) {
  SHINEON_REQUIRE(a && b && y && n > 0, "add_nhwc: bad argument");
  const bool al = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  const long n4 = al ? n / 4 : 0;
  klaunch(add_nhwc_kernel, grid_1d(n4 > 0 ? n4 : 1, 256), 256, 0, (cudaStream_t)stream, a, b, y, n4, n);
  return after_launch("add_nhwc_kernel");
}
