// GMM glue between the conv stacks (reference: models/networks/cpvton/warp.py):
//   * FeatureL2Norm x2 + FeatureCorrelation (warp.py:39-67) fused: an fp32 tiled GEMM per image
//     out[b, pB, iA] = <A[pA], B[pB]> / (|A[pA]| |B[pB]|),  iA = wA*h + hA  (column-major A index, warp.py:60)
//   * FeatureRegression tail: flatten (NCHW order) -> Linear -> tanh (warp.py:94-99)
#include "common.cuh"

namespace shineon {

constexpr int kTA = 64;  // pA (output channel) tile
constexpr int kTB = 64;  // pB (output pixel) tile
constexpr int kTK = 32;

// 128 threads, 8 pA x 4 pB accumulators per thread: per k three 128-bit shared loads feed 32 product FMAs (+ 12 for the
// squared norms).  The first version (4 x 2 per thread, scalar shared loads) ran at 17 TFLOP/s, shared-memory bound.
__global__ void __launch_bounds__(128)
    l2norm_corr_kernel(const float* __restrict__ fA, const float* __restrict__ fB, float* __restrict__ corr,
                       plane_t* __restrict__ yh, plane_t* __restrict__ yl, int h, int w, int C, int cpad, int fmt,
                       int normalize) {
  pdl_grid_sync();
  __shared__ __align__(16) float sA[kTK][kTA + 4];
  __shared__ __align__(16) float sB[kTK][kTB + 4];
  const int P = h * w;
  const int b = blockIdx.z;
  const int a0 = blockIdx.x * kTA, b0 = blockIdx.y * kTB;
  const int tid = threadIdx.x;
  const int tx = tid % 8, ty = tid / 8;  // pA = {4tx..4tx+3} U {32+4tx..32+4tx+3}, pB = 4ty..4ty+3
  const float* A = fA + (long)b * P * C;
  const float* B = fB + (long)b * P * C;
  float acc[4][8] = {};
  float na[8] = {}, nb[4] = {};
  for (int k0 = 0; k0 < C; k0 += kTK) {
    for (int idx = tid; idx < kTA * (kTK / 4); idx += 128) {
      const int row = idx / (kTK / 4), kq = idx % (kTK / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a0 + row < P && k0 + kq * 4 < C) v = *reinterpret_cast<const float4*>(A + (long)(a0 + row) * C + k0 + kq * 4);
      sA[kq * 4 + 0][row] = v.x; sA[kq * 4 + 1][row] = v.y; sA[kq * 4 + 2][row] = v.z; sA[kq * 4 + 3][row] = v.w;
    }
    for (int idx = tid; idx < kTB * (kTK / 4); idx += 128) {
      const int row = idx / (kTK / 4), kq = idx % (kTK / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b0 + row < P && k0 + kq * 4 < C) v = *reinterpret_cast<const float4*>(B + (long)(b0 + row) * C + k0 + kq * 4);
      sB[kq * 4 + 0][row] = v.x; sB[kq * 4 + 1][row] = v.y; sB[kq * 4 + 2][row] = v.z; sB[kq * 4 + 3][row] = v.w;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kTK; ++k) {
      const float4 a_lo = *reinterpret_cast<const float4*>(&sA[k][tx * 4]);
      const float4 a_hi = *reinterpret_cast<const float4*>(&sA[k][32 + tx * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&sB[k][ty * 4]);
      const float av[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) na[i] = fmaf(av[i], av[i], na[i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        nb[j] = fmaf(bv[j], bv[j], nb[j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(av[i], bv[j], acc[j][i]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int pB = b0 + ty * 4 + j;
    if (pB >= P) continue;
    const float inb = normalize ? 1.f / sqrtf(nb[j] + 1e-6f) : 1.f;  // warp.py:44-49
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int pA = a0 + (i < 4 ? tx * 4 + i : 32 + tx * 4 + (i - 4));
      if (pA >= P) continue;
      const float ina = normalize ? 1.f / sqrtf(na[i] + 1e-6f) : 1.f;
      const float v = acc[j][i] * ina * inb;
      const int hA = pA / w, wA = pA - hA * w;
      const int iA = wA * h + hA;  // feature_A.transpose(2,3) (warp.py:60)
      if (corr) corr[((long)b * P + pB) * P + iA] = v;
      if (yh) {
        plane_t hh, ll;
        split16(v, fmt, hh, ll);
        const long o = ((long)b * P + pB) * cpad + iA;
        yh[o] = hh;
        if (yl) yl[o] = ll;
      }
    }
  }
}

// FeatureL2Norm of both feature maps written as the operands of a per-image tensor-core GEMM (conv_igemm, w_per_image):
//   B' [b][pB][c]          = scale * fB[b,pB,c] / |fB[b,pB,:]|        the "activation" planes, pixel order hB*w + wB
//   A' [b][iA = wA*h+hA][c] = scale * fA[b,pA,c] / |fA[b,pA,:]|        the per-image "weights", rows in the transposed
//                                                                      pixel order of warp.py:60 (feature_A.transpose(2,3))
// so that out[b,pB,iA] = <B'[pB], A'[iA]> / scale^2 is FeatureCorrelation's output.  scale (a power of two) keeps the lo
// halves of the ~1/sqrt(C)-sized components out of the fp16 subnormals.  One warp per pixel.
__global__ void __launch_bounds__(256)
    l2norm_planes_kernel(const float* __restrict__ fA, const float* __restrict__ fB, plane_t* __restrict__ ah,
                         plane_t* __restrict__ al, plane_t* __restrict__ bh, plane_t* __restrict__ bl, int h, int w, int C,
                         int fmt, float scale) {
  pdl_grid_sync();
  const int P = h * w;
  const int b = blockIdx.z, isA = blockIdx.y;
  const int p = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (p >= P) return;
  const float* src = (isA ? fA : fB) + ((long)b * P + p) * C;
  float ss = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + c));
    ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
  }
  ss = warp_sum(ss);
  const float inv = scale / sqrtf(ss + 1e-6f);  // warp.py:44-49
  const int hp = p / w, wp = p - hp * w;
  const long row = (long)b * P + (isA ? wp * h + hp : p);
  plane_t* dh = (isA ? ah : bh) + row * C;
  plane_t* dl = (isA ? al : bl);
  if (dl) dl += row * C;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + c));
    plane_t hh[4], ll[4];
    split16(v.x * inv, fmt, hh[0], ll[0]);
    split16(v.y * inv, fmt, hh[1], ll[1]);
    split16(v.z * inv, fmt, hh[2], ll[2]);
    split16(v.w * inv, fmt, hh[3], ll[3]);
    *reinterpret_cast<uint2*>(dh + c) = make_uint2((uint32_t)hh[0] | ((uint32_t)hh[1] << 16), (uint32_t)hh[2] | ((uint32_t)hh[3] << 16));
    if (dl) *reinterpret_cast<uint2*>(dl + c) = make_uint2((uint32_t)ll[0] | ((uint32_t)ll[1] << 16), (uint32_t)ll[2] | ((uint32_t)ll[3] << 16));
  }
}

// FeatureL2Norm.forward on the reference layout (warp.py:43-50): y[b,c,p] = x[b,c,p] / sqrt(sum_c x[b,c,p]^2 + 1e-6).
// One thread per pixel, channel planes read coalesced across the warp.
__global__ void __launch_bounds__(128)
    feature_l2norm_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  const int p = blockIdx.x * 128 + threadIdx.x;
  if (p >= HW) return;
  const float* xb = x + (long)b * C * HW + p;
  float s = 0.f;
  for (int c = 0; c < C; ++c) {
    const float v = __ldg(xb + (long)c * HW);
    s = fmaf(v, v, s);
  }
  const float inv = 1.f / sqrtf(s + 1e-6f);
  float* yb = y + (long)b * C * HW + p;
  for (int c = 0; c < C; ++c) yb[(long)c * HW] = __ldg(xb + (long)c * HW) * inv;
}

// theta[b, o] = tanh(bias[o] + sum_{c,y,x} W[o, c*h*w + y*w + x] * x_nhwc[b, y, x, c]); one warp per output.
__global__ void __launch_bounds__(256)
    linear_tanh_kernel(const float* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias,
                       float* __restrict__ theta, int hw, int C, int out_dim) {
  pdl_grid_sync();
  const int b = blockIdx.x;
  const int K = hw * C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* xb = x + (long)b * K;
  for (int o = warp; o < out_dim; o += nw) {
    const float* wr = wgt + (long)o * K;
    float acc = 0.f;
    for (int e = lane; e < K; e += 32) {  // e indexes the NHWC activation (coalesced); map to the NCHW weight column
      const int c = e % C, p = e / C;
      acc = fmaf(xb[e], __ldg(wr + (long)c * hw + p), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) theta[(long)b * out_dim + o] = tanhf(acc + (bias ? bias[o] : 0.f));
  }
}

}  // namespace shineon

using namespace shineon;

extern "C" int shineon_l2norm_correlation(const float* featA, const float* featB, float* corr_f32, void* y_hi,
                                          void* y_lo, int B, int h, int w, int C, int cpad, int plane_fmt,
                                          int normalize, shineon_stream_t stream) {
  SHINEON_REQUIRE(plane_fmt == SHINEON_FMT_BF16 || plane_fmt == SHINEON_FMT_FP16, "l2norm_correlation: plane_fmt %d", plane_fmt);
  SHINEON_REQUIRE(featA && featB && (corr_f32 || y_hi), "l2norm_correlation: null pointer");
  SHINEON_REQUIRE(B > 0 && B <= 65535 && h > 0 && w > 0 && C > 0 && C % 4 == 0, "l2norm_correlation: bad shape (C %% 4)");
  SHINEON_REQUIRE(!y_hi || cpad >= h * w, "l2norm_correlation: cpad < h*w");
  const int P = h * w;
  dim3 grid(cdiv(P, kTA), cdiv(P, kTB), B);
  klaunch(l2norm_corr_kernel, grid, 128, 0, (cudaStream_t)stream, featA, featB, corr_f32, (plane_t*)y_hi,
                                                           (plane_t*)y_lo, h, w, C, cpad, plane_fmt, normalize);
  return after_launch("l2norm_corr_kernel");
}

extern "C" int shineon_feature_l2norm(const float* x, float* y, int B, int C, int H, int W, shineon_stream_t stream) {
  SHINEON_REQUIRE(x && y, "feature_l2norm: null pointer");
  SHINEON_REQUIRE(B > 0 && B <= 65535 && C > 0 && H > 0 && W > 0, "feature_l2norm: bad shape");
  klaunch(feature_l2norm_kernel, dim3(cdiv(H * W, 128), B), 128, 0, (cudaStream_t)stream, x, y, C, H * W);
  return after_launch("feature_l2norm_kernel");
}

extern "C" int shineon_linear_tanh(const float* x, const float* weight, const float* bias, float* theta, int B, int h,
                                   int w, int C, int out_dim, shineon_stream_t stream) {
  SHINEON_REQUIRE(x && weight && theta, "linear_tanh: null pointer");
  SHINEON_REQUIRE(B > 0 && h > 0 && w > 0 && C > 0 && out_dim > 0, "linear_tanh: bad shape");
  klaunch(linear_tanh_kernel, B, 256, 0, (cudaStream_t)stream, x, weight, bias, theta, h * w, C, out_dim);
  return after_launch("linear_tanh_kernel");
}

extern "C" int shineon_l2norm_planes(const float* featA, const float* featB, void* a_hi, void* a_lo, void* b_hi, void* b_lo,
                                     int B, int h, int w, int C, int plane_fmt, float scale, shineon_stream_t stream) {
  SHINEON_REQUIRE(plane_fmt == SHINEON_FMT_BF16 || plane_fmt == SHINEON_FMT_FP16, "l2norm_planes: plane_fmt %d", plane_fmt);
  SHINEON_REQUIRE(featA && featB && a_hi && b_hi && (a_lo == nullptr) == (b_lo == nullptr), "l2norm_planes: null pointer");
  SHINEON_REQUIRE(B > 0 && B <= 65535 && h > 0 && w > 0 && C > 0 && C % 64 == 0, "l2norm_planes: bad shape (C %% 64)");
  SHINEON_REQUIRE(scale > 0.f, "l2norm_planes: scale");
  klaunch(l2norm_planes_kernel, dim3(cdiv(h * w, 8), 2, B), 256, 0, (cudaStream_t)stream, featA, featB, (plane_t*)a_hi, (plane_t*)a_lo,
                                                                                    (plane_t*)b_hi, (plane_t*)b_lo, h, w, C,
                                                                                    plane_fmt, scale);
  return after_launch("l2norm_planes_kernel");
}
