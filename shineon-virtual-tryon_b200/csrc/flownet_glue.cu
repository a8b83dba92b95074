// FlowNet2 glue between the sub-networks (reference: models/flownet2_pytorch/models.py:127-192 and
// models/flownet.py:42-63): rgb-mean normalisation, x4 flow up-sampling, the "warp + difference + norm + concat"
// inputs of FlowNetS / FlowNetFusion, and the confidence mask.  The reference runs these as ~25 separate
// elementwise / Resample2d / ChannelNorm launches; here each concat input is produced by one kernel.
#include "common.cuh"

namespace shineon {

// Resample2d forward taps (resample2d_kernel.cu:37-59), shared by the fused concat kernels.
struct WarpTap {
  int o00, o01, o10, o11;
  float w00, w01, w10, w11;
};
__device__ __forceinline__ WarpTap warp_tap(int x, int y, float dx, float dy, int H, int W) {
  const float xf = (float)x + dx, yf = (float)y + dy;
  const float fx = floorf(xf), fy = floorf(yf);
  const float alpha = xf - fx, beta = yf - fy;
  const float cfx = fminf(fmaxf(fx, -4.f), (float)W + 4.f), cfy = fminf(fmaxf(fy, -4.f), (float)H + 4.f);
  const int xL = max(min((int)cfx, W - 1), 0), xR = max(min((int)cfx + 1, W - 1), 0);
  const int yT = max(min((int)cfy, H - 1), 0), yB = max(min((int)cfy + 1, H - 1), 0);
  WarpTap t;
  t.o00 = yT * W + xL; t.o01 = yT * W + xR; t.o10 = yB * W + xL; t.o11 = yB * W + xR;
  t.w00 = (1.f - alpha) * (1.f - beta); t.w01 = alpha * (1.f - beta);
  t.w10 = (1.f - alpha) * beta; t.w11 = alpha * beta;
  return t;
}
__device__ __forceinline__ float warp_sample(const float* __restrict__ pl, const WarpTap& t) {
  float v = 0.f;
  v += t.w00 * __ldg(pl + t.o00);
  v += t.w01 * __ldg(pl + t.o01);
  v += t.w10 * __ldg(pl + t.o10);
  v += t.w11 * __ldg(pl + t.o11);
  return v;
}

// ---- rgb mean over (frame, H, W) per (b, c)  (models.py:128)
__global__ void __launch_bounds__(256)
    flownet_mean_kernel(const float* __restrict__ in, double* __restrict__ ws, long per_bc) {
  pdl_grid_sync();
  const int bc = blockIdx.y;
  const float* p = in + (long)bc * per_bc;
  float s = 0.f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < per_bc; i += (long)gridDim.x * blockDim.x) s += p[i];
  s = warp_sum(s);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(ws + bc, (double)t);
  }
}
// x[b, f*3+c, y, x] = (in[b, c, f, y, x] - mean[b,c]) / rgb_max   (models.py:130-133)
__global__ void __launch_bounds__(256)
    flownet_normalize_kernel(const float* __restrict__ in, const double* __restrict__ ws, float* __restrict__ x, int HW,
                             float rgb_max, long total) {
  pdl_grid_sync();
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int p = (int)(e % HW);
    const int ch = (int)((e / HW) % 6);
    const int b = (int)(e / ((long)HW * 6));
    const int f = ch / 3, c = ch - f * 3;
    const float mean = (float)(ws[b * 3 + c] / (2.0 * HW));
    x[e] = (in[(((long)b * 3 + c) * 2 + f) * HW + p] - mean) / rgb_max;
  }
}

// ---- nn.Upsample(scale_factor=4) of a 2-channel flow held as f32 NHWC [B,h,w,cs]; dst NCHW [B,2,4h,4w]
//      value = up(src * mul)  (models.py:137,149,161,170)
__global__ void __launch_bounds__(256)
    upsample4x_flow_kernel(const float* __restrict__ src, int cs, float* __restrict__ dst, int h, int w, float mul,
                           int bilinear, long total) {
  pdl_grid_sync();
  const int H = 4 * h, W = 4 * w;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(e % W), oy = (int)((e / W) % H);
    const int c = (int)((e / ((long)W * H)) % 2), b = (int)(e / ((long)W * H * 2));
    const float* s = src + (long)b * h * w * cs + c;
    float v;
    if (bilinear) {  // align_corners=False: src = max(0.25*(dst+0.5)-0.5, 0)
      float sy = fmaxf(0.25f * ((float)oy + 0.5f) - 0.5f, 0.f), sx = fmaxf(0.25f * ((float)ox + 0.5f) - 0.5f, 0.f);
      const int y0 = (int)sy, x0 = (int)sx;
      const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
      const float ly1 = sy - (float)y0, lx1 = sx - (float)x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
      const float v00 = s[((long)y0 * w + x0) * cs] * mul, v01 = s[((long)y0 * w + x1) * cs] * mul;
      const float v10 = s[((long)y1 * w + x0) * cs] * mul, v11 = s[((long)y1 * w + x1) * cs] * mul;
      v = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
    } else {  // nearest: src = floor(dst / 4)
      v = s[((long)(oy >> 2) * w + (ox >> 2)) * cs] * mul;
    }
    dst[e] = v;
  }
}

// ---- concat1 / concat2 (models.py:139-146, 151-158): [x(6) | resample(x[3:6], flow)(3) | flow/div(2) | ||x[:3]-res||(1)]
__global__ void __launch_bounds__(256)
    flownet_warp_concat_kernel(const float* __restrict__ x, const float* __restrict__ flow, float* __restrict__ out,
                               int H, int W, float div_flow) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  const int HW = H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const int y = p / W, xx = p - y * W;
    const float* xb = x + (long)b * 6 * HW;
    float* ob = out + (long)b * 12 * HW;
    const float dx = flow[((long)b * 2 + 0) * HW + p], dy = flow[((long)b * 2 + 1) * HW + p];
    const WarpTap t = warp_tap(xx, y, dx, dy, H, W);
    float nrm = 0.f;
#pragma unroll
    for (int c = 0; c < 6; ++c) ob[(long)c * HW + p] = xb[(long)c * HW + p];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float r = warp_sample(xb + (long)(3 + c) * HW, t);
      ob[(long)(6 + c) * HW + p] = r;
      const float d = xb[(long)c * HW + p] - r;
      nrm += d * d;
    }
    ob[(long)9 * HW + p] = dx / div_flow;
    ob[(long)10 * HW + p] = dy / div_flow;
    ob[(long)11 * HW + p] = sqrtf(nrm);
  }
}

// ---- concat3 (models.py:160-185): [x[:3] | flow_sd(2) | flow_s2(2) | |flow_sd| | |flow_s2| | ||x[:3]-res_sd|| | ||x[:3]-res_s2||]
__global__ void __launch_bounds__(256)
    flownet_fusion_concat_kernel(const float* __restrict__ x, const float* __restrict__ fsd, const float* __restrict__ fs2,
                                 float* __restrict__ out, int H, int W) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  const int HW = H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const int y = p / W, xx = p - y * W;
    const float* xb = x + (long)b * 6 * HW;
    float* ob = out + (long)b * 11 * HW;
    const float sdx = fsd[((long)b * 2) * HW + p], sdy = fsd[((long)b * 2 + 1) * HW + p];
    const float s2x = fs2[((long)b * 2) * HW + p], s2y = fs2[((long)b * 2 + 1) * HW + p];
    const WarpTap tsd = warp_tap(xx, y, sdx, sdy, H, W), ts2 = warp_tap(xx, y, s2x, s2y, H, W);
    float nsd = 0.f, ns2 = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x0 = xb[(long)c * HW + p];
      ob[(long)c * HW + p] = x0;
      const float d1 = x0 - warp_sample(xb + (long)(3 + c) * HW, tsd);
      const float d2 = x0 - warp_sample(xb + (long)(3 + c) * HW, ts2);
      nsd += d1 * d1;
      ns2 += d2 * d2;
    }
    ob[(long)3 * HW + p] = sdx; ob[(long)4 * HW + p] = sdy;
    ob[(long)5 * HW + p] = s2x; ob[(long)6 * HW + p] = s2y;
    ob[(long)7 * HW + p] = sqrtf(sdx * sdx + sdy * sdy);
    ob[(long)8 * HW + p] = sqrtf(s2x * s2x + s2y * s2y);
    ob[(long)9 * HW + p] = sqrtf(nsd);
    ob[(long)10 * HW + p] = sqrtf(ns2);
  }
}

// ---- conf = (sum_c (im1 - resample(im2, flow))^2 < thr)  (models/flownet.py:55,61-62)
__global__ void __launch_bounds__(256)
    flow_confidence_kernel(const float* __restrict__ im1, const float* __restrict__ im2, const float* __restrict__ flow,
                           float* __restrict__ conf, int C, int H, int W, float thr) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  const int HW = H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const int y = p / W, xx = p - y * W;
    const WarpTap t = warp_tap(xx, y, flow[((long)b * 2) * HW + p], flow[((long)b * 2 + 1) * HW + p], H, W);
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
      const float d = im1[((long)b * C + c) * HW + p] - warp_sample(im2 + ((long)b * C + c) * HW, t);
      s += d * d;
    }
    conf[(long)b * HW + p] = s < thr ? 1.f : 0.f;
  }
}

// ---- nn.Upsample(size=(Ho, Wo), mode="bilinear") (align_corners=False), models/flownet.py:47-51,56-58:
// ATen upsample_bilinear2d: scale = in / out (float), src = max(scale * (dst + 0.5) - 0.5, 0), i1 = i0 + (i0 < in - 1),
// value = l0y * (l0x * p00 + l1x * p01) + l1y * (l0x * p10 + l1x * p11), times `mul` (the flow rescale old_h / new_h).
__global__ void __launch_bounds__(256)
    bilinear_resize_kernel(const float* __restrict__ x, float* __restrict__ y, int Hi, int Wi, int Ho, int Wo, float sy,
                           float sx, float mul, long total) {
  pdl_grid_sync();
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(e % Wo), oy = (int)((e / Wo) % Ho);
    const long bc = e / ((long)Wo * Ho);
    float fy = __fsub_rn(__fmul_rn(sy, (float)oy + 0.5f), 0.5f), fx = __fsub_rn(__fmul_rn(sx, (float)ox + 0.5f), 0.5f);
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < Hi - 1 ? 1 : 0), x1 = x0 + (x0 < Wi - 1 ? 1 : 0);
    const float ly1 = fy - (float)y0, lx1 = fx - (float)x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const float* p = x + bc * Hi * Wi;
    const float v = ly0 * (lx0 * __ldg(p + (long)y0 * Wi + x0) + lx1 * __ldg(p + (long)y0 * Wi + x1)) +
                    ly1 * (lx0 * __ldg(p + (long)y1 * Wi + x0) + lx1 * __ldg(p + (long)y1 * Wi + x1));
    y[e] = v * mul;
  }
}

// `upsampled_flow*`: ConvTranspose2d(2, 2, 4, 2, 1) on a predicted flow (FlowNetC.py:59-62, FlowNetS / SD / Fusion alike).
// Two channels in, two out: 16 MACs per output value -- on the tensor-core conv this was a launch over a 64-channel padded
// operand plus a f32 -> planes conversion of the flow in front of it, 18 times per FlowNet2 forward.  One thread per output
// pixel reads the (at most) 2 x 2 contributing flow vectors and writes both channels of the consumer's concat window as
// 16-bit hi/lo planes.  out[2*iy - 1 + ky][2*ix - 1 + kx][co] += in[iy][ix][ci] * w[ci][co][ky][kx].
template <int FMT>
__global__ void __launch_bounds__(256)
    flow_deconv4x4s2_planes_kernel(const float* __restrict__ flow, int flow_cstride, const float* __restrict__ weight,
                                   const float* __restrict__ bias, plane_t* __restrict__ yh, plane_t* __restrict__ yl,
                                   int cstride, int h, int w, long total) {
  pdl_grid_sync();
  __shared__ float s_w[64];  // [ci][co][ky][kx]
  if (threadIdx.x < 64) s_w[threadIdx.x] = __ldg(weight + threadIdx.x);
  __syncthreads();
  const float b0 = bias ? __ldg(bias) : 0.f, b1 = bias ? __ldg(bias + 1) : 0.f;
  const int W2 = 2 * w, H2 = 2 * h;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(e % W2), oy = (int)((e / W2) % H2);
    const long b = e / ((long)W2 * H2);
    float a0 = b0, a1 = b1;
    const int iy_hi = (oy + 1) >> 1, ix_hi = (ox + 1) >> 1;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int iy = iy_hi - dy, ky = oy + 1 - 2 * iy;  // ky in {0,1} + 2*dy
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int ix = ix_hi - dx, kx = ox + 1 - 2 * ix;
        if (ix < 0 || ix >= w) continue;
        const float* src = flow + ((b * h + iy) * w + ix) * flow_cstride;
        const float v0 = __ldg(src), v1 = __ldg(src + 1);
        const int t = ky * 4 + kx;
        a0 = fmaf(v0, s_w[t], a0);            // ci 0, co 0
        a0 = fmaf(v1, s_w[32 + t], a0);       // ci 1, co 0
        a1 = fmaf(v0, s_w[16 + t], a1);       // ci 0, co 1
        a1 = fmaf(v1, s_w[48 + t], a1);       // ci 1, co 1
      }
    }
    uint32_t hi, lo;
    split16x2(a0, a1, FMT, hi, lo);
    *reinterpret_cast<uint32_t*>(yh + e * cstride) = hi;
    if (yl) *reinterpret_cast<uint32_t*>(yl + e * cstride) = lo;
  }
}

static inline int blocks_for(long total) {
  long b = (total + 255) / 256;
  return (int)(b > 148 * 32 ? 148 * 32 : (b < 1 ? 1 : b));
}

}  // namespace shineon

using namespace shineon;

extern "C" int shineon_flownet_normalize(const float* inputs, float* x, double* ws, int B, int H, int W, float rgb_max,
                                         shineon_stream_t stream_) {
  SHINEON_REQUIRE(inputs && x && ws, "flownet_normalize: null pointer");
  SHINEON_REQUIRE(B > 0 && B * 3 <= 65535 && H > 0 && W > 0 && rgb_max != 0.f, "flownet_normalize: bad shape");
  cudaStream_t stream = (cudaStream_t)stream_;
  cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(double) * 3 * (size_t)B, stream);
  if (e != cudaSuccess) return fail(SHINEON_ERR_CUDA, "flownet_normalize memset: %s", cudaGetErrorString(e));
  const long per_bc = 2l * H * W;
  klaunch(flownet_mean_kernel, dim3(blocks_for(per_bc) > 64 ? 64 : blocks_for(per_bc), B * 3), 256, 0, stream, inputs, ws, per_bc);
  int rc = after_launch("flownet_mean_kernel");
  if (rc) return rc;
  const long total = (long)B * 6 * H * W;
  klaunch(flownet_normalize_kernel, blocks_for(total), 256, 0, stream, inputs, ws, x, H * W, rgb_max, total);
  return after_launch("flownet_normalize_kernel");
}

extern "C" int shineon_upsample4x_flow(const float* src, int src_cstride, float* dst, int B, int h, int w, float mul,
                                       int bilinear, shineon_stream_t stream) {
  SHINEON_REQUIRE(src && dst && src_cstride >= 2, "upsample4x_flow: bad argument");
  SHINEON_REQUIRE(B > 0 && h > 0 && w > 0, "upsample4x_flow: bad shape");
  const long total = (long)B * 2 * 16 * h * w;
  klaunch(upsample4x_flow_kernel, blocks_for(total), 256, 0, (cudaStream_t)stream, src, src_cstride, dst, h, w, mul, bilinear, total);
  return after_launch("upsample4x_flow_kernel");
}

extern "C" int shineon_flownet_warp_concat(const float* x, const float* flow, float* out, int B, int H, int W,
                                           float div_flow, shineon_stream_t stream) {
  SHINEON_REQUIRE(x && flow && out && div_flow != 0.f, "flownet_warp_concat: bad argument");
  SHINEON_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, "flownet_warp_concat: bad shape");
  klaunch(flownet_warp_concat_kernel, dim3(blocks_for((long)H * W), B), 256, 0, (cudaStream_t)stream, x, flow, out, H, W, div_flow);
  return after_launch("flownet_warp_concat_kernel");
}

extern "C" int shineon_flownet_fusion_concat(const float* x, const float* flow_sd, const float* flow_s2, float* out, int B,
                                             int H, int W, shineon_stream_t stream) {
  SHINEON_REQUIRE(x && flow_sd && flow_s2 && out, "flownet_fusion_concat: null pointer");
  SHINEON_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, "flownet_fusion_concat: bad shape");
  klaunch(flownet_fusion_concat_kernel, dim3(blocks_for((long)H * W), B), 256, 0, (cudaStream_t)stream, x, flow_sd, flow_s2, out, H, W);
  return after_launch("flownet_fusion_concat_kernel");
}

extern "C" int shineon_flow_confidence(const float* im1, const float* im2, const float* flow, float* conf, int B, int C,
                                       int H, int W, float threshold, shineon_stream_t stream) {
  SHINEON_REQUIRE(im1 && im2 && flow && conf, "flow_confidence: null pointer");
  SHINEON_REQUIRE(B > 0 && B <= 65535 && C > 0 && H > 0 && W > 0, "flow_confidence: bad shape");
  klaunch(flow_confidence_kernel, dim3(blocks_for((long)H * W), B), 256, 0, (cudaStream_t)stream, im1, im2, flow, conf, C, H, W, threshold);
  return after_launch("flow_confidence_kernel");
}

extern "C" int shineon_bilinear_resize(const float* x, float* y, int BC, int Hi, int Wi, int Ho, int Wo, float mul,
                                       shineon_stream_t stream) {
  SHINEON_REQUIRE(x && y, "bilinear_resize: null pointer");
  SHINEON_REQUIRE(BC > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "bilinear_resize: bad shape");
  const long total = (long)BC * Ho * Wo;
  klaunch(bilinear_resize_kernel, blocks_for(total), 256, 0, (cudaStream_t)stream, x, y, Hi, Wi, Ho, Wo, (float)Hi / (float)Ho,
                                                                            (float)Wi / (float)Wo, mul, total);
  return after_launch("bilinear_resize_kernel");
}

extern "C" int shineon_flow_deconv4x4s2_planes(const float* flow, int flow_cstride, const float* weight, const float* bias,
                                               void* y_hi, void* y_lo, int y_cstride, int B, int h, int w, int plane_fmt,
                                               shineon_stream_t stream) {
  SHINEON_REQUIRE(flow && weight && y_hi && flow_cstride >= 2, "flow_deconv4x4s2_planes: bad argument");
  SHINEON_REQUIRE(B > 0 && h > 0 && w > 0 && y_cstride >= 2 && y_cstride % 2 == 0, "flow_deconv4x4s2_planes: bad shape");
  SHINEON_REQUIRE(plane_fmt == SHINEON_FMT_BF16 || plane_fmt == SHINEON_FMT_FP16, "flow_deconv4x4s2_planes: plane_fmt %d", plane_fmt);
  SHINEON_REQUIRE((reinterpret_cast<uintptr_t>(y_hi) & 3) == 0 && (reinterpret_cast<uintptr_t>(y_lo) & 3) == 0, "flow_deconv4x4s2_planes: planes must be 4-byte aligned");
  const long total = (long)B * 4 * h * w;
  if (plane_fmt == SHINEON_FMT_FP16)
    klaunch(flow_deconv4x4s2_planes_kernel<SHINEON_FMT_FP16>, blocks_for(total), 256, 0, (cudaStream_t)stream, flow, flow_cstride, weight,
            bias, (plane_t*)y_hi, (plane_t*)y_lo, y_cstride, h, w, total);
  else
    klaunch(flow_deconv4x4s2_planes_kernel<SHINEON_FMT_BF16>, blocks_for(total), 256, 0, (cudaStream_t)stream, flow, flow_cstride, weight,
            bias, (plane_t*)y_hi, (plane_t*)y_lo, y_cstride, h, w, total);
  return after_launch("flow_deconv4x4s2_planes_kernel");
}
