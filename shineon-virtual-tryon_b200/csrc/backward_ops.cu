// Memory-bound backward passes of the U-Net / TOM training step (the reference gets them from
// autograd over models/networks/cpvton/unet.py:129-198, attention/sagan.py:29-53, models/unet_mask_model.py:64-217):
//   * instnorm_act_bwd     : d(act o InstanceNorm2d)  -> f32 and/or 16-bit planes (the dgrad / wgrad operand)
//   * act_bwd              : g * act'(z)
//   * upsample2x_cat_bwd   : adjoint of bilinear x2 (align_corners=False) + channel split of the concat
//   * sagan_attention_bwd  : dq, dk, dv, dgamma of softmax(q^T k) attention (N <= 192 tokens)
//   * tom_compose_bwd      : tanh / sigmoid / mask compose (+ flow-warp blend) backward
//   * l1_loss              : mean |a-b| and its gradient (image, mask and VGG-feature terms)
//   * maxpool2x2 fwd/bwd, relu mask (VGG19 perceptual loss, models/networks/vgg.py:6-38)
// All tensors NHWC f32 unless noted.
#include "common.cuh"

namespace shineon {

__device__ __forceinline__ float act_grad(float z, int act, float param) {
  switch (act) {
    case SHINEON_ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case SHINEON_ACT_LEAKY: return z > 0.f ? 1.f : param;
    case SHINEON_ACT_GELU:
      return 0.5f * (1.f + erff(z * 0.70710678118654752440f)) + z * 0.39894228040143267794f * expf(-0.5f * z * z);
    case SHINEON_ACT_SWISH: {
      const float s = 1.f / (1.f + expf(-z));
      return s * (1.f + z * (1.f - s));
    }
    case SHINEON_ACT_SINE: return 30.f * cosf(30.f * z);
    case SHINEON_ACT_TANH: {
      const float t = tanhf(z);
      return 1.f - t * t;
    }
    case SHINEON_ACT_SIGMOID: {
      const float s = 1.f / (1.f + expf(-z));
      return s * (1.f - s);
    }
    default: return 1.f;
  }
}

static inline int grid_x(long total, int threads) {
  long b = (total + threads - 1) / threads;
  return (int)(b > 148 * 32 ? 148 * 32 : (b < 1 ? 1 : b));
}

// ------------------------------------------------------------------------------ instance-norm backward
// forward: yhat = (x - mean) * rstd ; a = act(yhat).   gz = (g1 + g2) * act'(yhat)
//   gx = rstd * (gz - mean_hw(gz) - yhat * mean_hw(gz * yhat))
// Pass 1: per-(n,c) S1 = sum gz, S2 = sum gz*yhat (fp32 partials per CTA, fp64 atomics across CTAs).
__global__ void __launch_bounds__(256)
    instnorm_bwd_stats_kernel(const float* __restrict__ x, const double* __restrict__ ws_fwd,
                              const float* __restrict__ g1, const float* __restrict__ g2, double* __restrict__ ws_bwd,
                              int HW, int C, int pix_per_cta, float eps, int act, float act_param) {
  pdl_grid_sync();
  __shared__ float s1[8][33], s2[8][33];
  const int n = blockIdx.z;
  const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + cl;
  const int p0 = blockIdx.x * pix_per_cta, p1 = min(p0 + pix_per_cta, HW);
  float a1 = 0.f, a2 = 0.f;
  if (c < C) {
    const double s = ws_fwd[((long)n * C + c) * 2], q = ws_fwd[((long)n * C + c) * 2 + 1];
    const double m = s / HW;
    double var = q / HW - m * m;
    if (var < 0.0) var = 0.0;
    const float mean = (float)m, rstd = (float)(1.0 / sqrt(var + (double)eps));
    const long base = (long)n * HW * C + c;
#pragma unroll 4
    for (int p = p0 + pl; p < p1; p += 8) {
      const long o = base + (long)p * C;
      const float yh = (__ldg(x + o) - mean) * rstd;
      float g = __ldg(g1 + o);
      if (g2) g += __ldg(g2 + o);
      g *= act_grad(yh, act, act_param);
      a1 += g;
      a2 = fmaf(g, yh, a2);
    }
  }
  s1[pl][cl] = a1;
  s2[pl][cl] = a2;
  __syncthreads();
  if (pl == 0 && c < C) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      t1 += s1[i][cl];
      t2 += s2[i][cl];
    }
    atomicAdd(ws_bwd + ((long)n * C + c) * 2 + 0, (double)t1);
    atomicAdd(ws_bwd + ((long)n * C + c) * 2 + 1, (double)t2);
  }
}

template <int VEC>
__global__ void __launch_bounds__(256)
    instnorm_bwd_apply_kernel(const float* __restrict__ x, const double* __restrict__ ws_fwd,
                              const double* __restrict__ ws_bwd, const float* __restrict__ g1,
                              const float* __restrict__ g2, float* __restrict__ gxf, plane_t* __restrict__ gxh,
                              plane_t* __restrict__ gxl, int HW, int C, int cpad, float eps, int do_norm, int act,
                              float act_param, int fmt) {
  pdl_grid_sync();
  extern __shared__ float s_tab[];  // mean[C], rstd[C], m1[C], m2[C]
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean = 0.f, rstd = 1.f, m1 = 0.f, m2 = 0.f;
    if (do_norm) {
      const double s = ws_fwd[((long)n * C + c) * 2], q = ws_fwd[((long)n * C + c) * 2 + 1];
      const double m = s / HW;
      double var = q / HW - m * m;
      if (var < 0.0) var = 0.0;
      mean = (float)m;
      rstd = (float)(1.0 / sqrt(var + (double)eps));
      m1 = (float)(ws_bwd[((long)n * C + c) * 2] / HW);
      m2 = (float)(ws_bwd[((long)n * C + c) * 2 + 1] / HW);
    }
    s_tab[c] = mean;
    s_tab[C + c] = rstd;
    s_tab[2 * C + c] = m1;
    s_tab[3 * C + c] = m2;
  }
  __syncthreads();
  const int cg = C / VEC;
  const unsigned total = (unsigned)HW * (unsigned)cg;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const unsigned p = e / (unsigned)cg;
    const int g = (int)(e - p * (unsigned)cg);
    const long xi = ((long)n * HW + p) * C + g * VEC;
    float xv[VEC], gv[VEC];
    if (VEC == 4) {
      const float4 t = *reinterpret_cast<const float4*>(x + xi);
      xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
      float4 u = *reinterpret_cast<const float4*>(g1 + xi);
      if (g2) {
        const float4 w = *reinterpret_cast<const float4*>(g2 + xi);
        u.x += w.x; u.y += w.y; u.z += w.z; u.w += w.w;
      }
      gv[0] = u.x; gv[1] = u.y; gv[2] = u.z; gv[3] = u.w;
    } else {
      xv[0] = x[xi];
      gv[0] = g1[xi] + (g2 ? g2[xi] : 0.f);
    }
    float out[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = g * VEC + j;
      const float yh = (xv[j] - s_tab[c]) * s_tab[C + c];
      const float gz = gv[j] * act_grad(yh, act, act_param);
      out[j] = do_norm ? s_tab[C + c] * (gz - s_tab[2 * C + c] - yh * s_tab[3 * C + c]) : gz;
    }
    if (gxf) {
      if (VEC == 4)
        *reinterpret_cast<float4*>(gxf + xi) = make_float4(out[0], out[1], out[2], out[3]);
      else
        gxf[xi] = out[0];
    }
    if (gxh) {
      const long po = ((long)n * HW + p) * cpad + g * VEC;
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        plane_t h, l;
        split16(out[j], fmt, h, l);
        gxh[po + j] = h;
        if (gxl) gxl[po + j] = l;
      }
    }
  }
}

// ------------------------------------------------------------------------------ pointwise activation backward
__global__ void __launch_bounds__(256)
    act_bwd_kernel(const float* __restrict__ z, const float* __restrict__ g1, const float* __restrict__ g2,
                   float* __restrict__ gz, long n, int act, float act_param) {
  pdl_grid_sync();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float g = g1[i];
    if (g2) g += g2[i];
    gz[i] = g * act_grad(z[i], act, act_param);
  }
}

// ------------------------------------------------------------------------------ upsample x2 + concat: backward
// Forward (align_corners=False, scale 2): out[2i] = .25 in[i-1] + .75 in[i], out[2i+1] = .75 in[i] + .25 in[i+1]
// with the neighbour index clamped into the image.  Adjoint: in[i] collects rows 2i-1..2i+2 with weights
// .25,.75,.75,.25; the clamped taps of the first/last output row fold back onto the border pixel.
__device__ __forceinline__ void up2_adj(int i, int n_in, int (&r)[4], float (&w)[4]) {
  r[0] = 2 * i - 1; r[1] = 2 * i; r[2] = 2 * i + 1; r[3] = 2 * i + 2;
  w[0] = 0.25f; w[1] = 0.75f; w[2] = 0.75f; w[3] = 0.25f;
  if (i == 0) { w[0] = 0.f; r[0] = 0; w[1] = 1.0f; }                    // out[0] = in[0]
  if (i == n_in - 1) { w[3] = 0.f; r[3] = 2 * i + 1; w[2] = 1.0f; }     // out[2n-1] = in[n-1]
}

__global__ void __launch_bounds__(256)
    upsample2x_cat_bwd_kernel(const float* __restrict__ gu, int cs_in, float* __restrict__ g0, int C0,
                              float* __restrict__ g1, int C1, int H, int W) {
  pdl_grid_sync();
  const int n = blockIdx.z, iy = blockIdx.y;
  const int Ct = C0 + C1;
  const int cg = (Ct + 3) / 4;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= W * cg) return;
  const int ix = t / cg, c = (t - ix * cg) * 4;
  int ry[4], rx[4];
  float wy[4], wx[4];
  up2_adj(iy, H, ry, wy);
  up2_adj(ix, W, rx, wx);
  const int Wo = 2 * W;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const bool vec = (cs_in & 3) == 0 && c + 3 < Ct;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (wy[a] == 0.f) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (wx[b] == 0.f) continue;
      const float w = wy[a] * wx[b];
      const float* src = gu + (((long)n * 2 * H + ry[a]) * Wo + rx[b]) * cs_in + c;
      if (vec) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src));
        acc[0] = fmaf(w, v.x, acc[0]); acc[1] = fmaf(w, v.y, acc[1]);
        acc[2] = fmaf(w, v.z, acc[2]); acc[3] = fmaf(w, v.w, acc[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < Ct) acc[j] = fmaf(w, __ldg(src + j), acc[j]);
      }
    }
  }
  const long pix = ((long)n * H + iy) * W + ix;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int cc = c + j;
    if (cc < C0) g0[pix * C0 + cc] = acc[j];
    else if (cc < Ct) g1[pix * C1 + (cc - C0)] = acc[j];
  }
}

// ------------------------------------------------------------------------------ SAGAN attention backward
// Pass A (one CTA per QT queries): recompute A = softmax(q k^T) rows, o = A v, D_i = <do_i, o_i>, dgamma += <g_i, o_i>,
//   dA_ij = <do_i, v_j>, dE_ij = A_ij (dA_ij - D_i), dq_i = sum_j dE_ij k_j; stores A and dE rows for pass B.
// Pass B (one CTA per JT keys): dk_j = sum_i dE_ij q_i, dv_j = sum_i A_ij do_i.
template <int QT>
__global__ void __launch_bounds__(128)
    sagan_bwd_a_kernel(const float* __restrict__ qkv, const float* __restrict__ gamma, const float* __restrict__ gout,
                       float* __restrict__ Abuf, float* __restrict__ dEbuf, float* __restrict__ gqkv,
                       double* __restrict__ ggamma, int HW, int C, int Cq) {
  pdl_grid_sync();
  extern __shared__ float sm[];  // q[QT][Cq] | e[QT][HW] | de[QT][HW] | red[QT][4] | D[QT]
  float* sq = sm;
  float* se = sq + QT * Cq;
  float* sde = se + QT * HW;
  float* sred = sde + QT * HW;
  float* sD = sred + QT * 4;
  const int n = blockIdx.y, i0 = blockIdx.x * QT;
  const int nq = min(QT, HW - i0);
  const int ld = 2 * Cq + C;
  const float* base = qkv + (long)n * HW * ld;
  const float* gbase = gout + (long)n * HW * C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float gm = __ldg(gamma);

  for (int e = tid; e < QT * Cq; e += 128) {
    const int q = e / Cq, c = e - q * Cq;
    sq[e] = q < nq ? base[(long)(i0 + q) * ld + c] : 0.f;
  }
  __syncthreads();
  for (int j = tid; j < HW; j += 128) {
    const float* kj = base + (long)j * ld + Cq;
    float acc[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) acc[q] = 0.f;
    for (int c = 0; c < Cq; ++c) {
      const float kv = kj[c];
#pragma unroll
      for (int q = 0; q < QT; ++q) acc[q] = fmaf(sq[q * Cq + c], kv, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < QT; ++q) se[q * HW + j] = acc[q];
  }
  __syncthreads();
  for (int q = warp; q < QT; q += 4) {
    float* e = se + q * HW;
    float m = -INFINITY;
    for (int j = lane; j < HW; j += 32) m = fmaxf(m, e[j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < HW; j += 32) {
      const float p = expf(e[j] - m);
      e[j] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < HW; j += 32) e[j] *= inv;
  }
  __syncthreads();
  // o[q][c] (thread owns channels c = tid, tid+128, ...), D_q = gamma * <g_q, o_q>, dgamma partial = <g_q, o_q>
  const float* vbase = base + 2 * Cq;
  float dpart[QT];
#pragma unroll
  for (int q = 0; q < QT; ++q) dpart[q] = 0.f;
  for (int c = tid; c < C; c += 128) {
    float acc[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) acc[q] = 0.f;
    for (int j = 0; j < HW; ++j) {
      const float v = vbase[(long)j * ld + c];
#pragma unroll
      for (int q = 0; q < QT; ++q) acc[q] = fmaf(se[q * HW + j], v, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < QT; ++q)
      if (q < nq) dpart[q] = fmaf(gbase[(long)(i0 + q) * C + c], acc[q], dpart[q]);
  }
#pragma unroll
  for (int q = 0; q < QT; ++q) {
    const float s = warp_sum(dpart[q]);
    if (lane == 0) sred[q * 4 + warp] = s;
  }
  __syncthreads();
  if (tid < QT) {
    const float go = sred[tid * 4] + sred[tid * 4 + 1] + sred[tid * 4 + 2] + sred[tid * 4 + 3];  // <g_q, o_q>
    sD[tid] = gm * go;
    if (tid < nq) atomicAdd(ggamma, (double)go);
  }
  __syncthreads();
  // dA[q][j] = gamma * <g_q, v_j>; one warp per key j, lanes over channels
  for (int j = warp; j < HW; j += 4) {
    float acc[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) acc[q] = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float v = vbase[(long)j * ld + c];
#pragma unroll
      for (int q = 0; q < QT; ++q)
        if (q < nq) acc[q] = fmaf(gbase[(long)(i0 + q) * C + c], v, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < QT; ++q) {
      const float s = warp_sum(acc[q]);
      if (lane == 0) sde[q * HW + j] = se[q * HW + j] * (gm * s - sD[q]);
    }
  }
  __syncthreads();
  for (int e = tid; e < QT * HW; e += 128) {
    const int q = e / HW, j = e - q * HW;
    if (q < nq) {
      Abuf[((long)n * HW + i0 + q) * HW + j] = se[e];
      dEbuf[((long)n * HW + i0 + q) * HW + j] = sde[e];
    }
  }
  // dq[q][c] = sum_j dE[q][j] k[j][c]
  for (int e = tid; e < QT * Cq; e += 128) {
    const int q = e / Cq, c = e - q * Cq;
    if (q >= nq) continue;
    float acc = 0.f;
    for (int j = 0; j < HW; ++j) acc = fmaf(sde[q * HW + j], base[(long)j * ld + Cq + c], acc);
    gqkv[((long)n * HW + i0 + q) * ld + c] = acc;
  }
}

__global__ void __launch_bounds__(128)
    sagan_bwd_b_kernel(const float* __restrict__ qkv, const float* __restrict__ gamma, const float* __restrict__ gout,
                       const float* __restrict__ Abuf, const float* __restrict__ dEbuf, float* __restrict__ gqkv,
                       int HW, int C, int Cq) {
  pdl_grid_sync();
  // one CTA per key j: dk_j[c'] = sum_i dE[i][j] q_i[c'],  dv_j[c] = gamma * sum_i A[i][j] g_i[c]
  extern __shared__ float sm[];  // a[HW] | de[HW]
  float* sa = sm;
  float* sde = sm + HW;
  const int n = blockIdx.y, j = blockIdx.x;
  const int ld = 2 * Cq + C;
  const float* base = qkv + (long)n * HW * ld;
  const float* gbase = gout + (long)n * HW * C;
  const float gm = __ldg(gamma);
  for (int i = threadIdx.x; i < HW; i += 128) {
    sa[i] = Abuf[((long)n * HW + i) * HW + j];
    sde[i] = dEbuf[((long)n * HW + i) * HW + j];
  }
  __syncthreads();
  float* dst = gqkv + ((long)n * HW + j) * ld;
  for (int c = threadIdx.x; c < Cq; c += 128) {
    float acc = 0.f;
    for (int i = 0; i < HW; ++i) acc = fmaf(sde[i], base[(long)i * ld + c], acc);
    dst[Cq + c] = acc;
  }
  for (int c = threadIdx.x; c < C; c += 128) {
    float acc = 0.f;
    for (int i = 0; i < HW; ++i) acc = fmaf(sa[i], gbase[(long)i * C + c], acc);
    dst[2 * Cq + c] = gm * acc;
  }
}

__global__ void scalar_finish_kernel(const double* __restrict__ acc, float* __restrict__ out, float alpha, float beta) {
  pdl_grid_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = beta == 0.f ? alpha * (float)acc[0] : fmaf(beta, out[0], alpha * (float)acc[0]);
}

// ------------------------------------------------------------------------------ TOM compose backward
__global__ void __launch_bounds__(256)
    tom_compose_bwd_kernel(const float* __restrict__ u, int Cout, const float* __restrict__ cloth,
                           const float* __restrict__ warped_prev, const float* __restrict__ g_rend,
                           const float* __restrict__ g_mask, const float* __restrict__ g_tryon,
                           const float* __restrict__ g_fmask, float* __restrict__ gu, float* __restrict__ g_warped,
                           int HW, int nf, int f, int flow_warp) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const float* up = u + ((long)b * HW + p) * Cout;
    float* gp = gu + ((long)b * HW + p) * Cout;
    const float m = 1.f / (1.f + expf(-up[3 * nf + f]));
    float fm = 0.f;
    if (flow_warp) fm = 1.f / (1.f + expf(-up[4 * nf + f]));
    // the gradients are per-frame tensors: g_rend / g_tryon [B,3,H,W], g_mask / g_fmask [B,1,H,W]
    float gm = g_mask ? g_mask[(long)b * HW + p] : 0.f;
    float gfm = (flow_warp && g_fmask) ? g_fmask[(long)b * HW + p] : 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const long o = ((long)b * 3 * nf + 3 * f + k) * HW + p;
      const float r = tanhf(up[3 * f + k]);
      const long og = ((long)b * 3 + k) * HW + p;
      const float gt = g_tryon ? g_tryon[og] : 0.f;
      const float grr = gt * (1.f - m);
      float rr = r, gr = grr;
      if (warped_prev) {
        const float w = warped_prev[((long)b * 3 + k) * HW + p];
        rr = (1.f - fm) * w + fm * r;
        gr = grr * fm;
        gfm = fmaf(grr, r - w, gfm);
        if (g_warped) g_warped[((long)b * 3 + k) * HW + p] = grr * (1.f - fm);
      }
      if (g_rend) gr += g_rend[og];
      gm = fmaf(gt, cloth[o] - rr, gm);
      gp[3 * f + k] = gr * (1.f - r * r);
    }
    gp[3 * nf + f] = gm * m * (1.f - m);
    if (flow_warp) gp[4 * nf + f] = gfm * fm * (1.f - fm);
  }
}

// ------------------------------------------------------------------------------ L1 loss (mean) + gradient
__global__ void __launch_bounds__(256)
    l1_loss_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ ga, double* __restrict__ acc,
                   long n, float gscale, int accumulate_grad) {
  pdl_grid_sync();
  float part = 0.f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    part += fabsf(d);
    if (ga) {
      const float g = d > 0.f ? gscale : (d < 0.f ? -gscale : 0.f);  // torch: sign(d)
      ga[i] = accumulate_grad ? ga[i] + g : g;
    }
  }
  part = warp_sum(part);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += s[i];
    atomicAdd(acc, (double)t);
  }
}

// ------------------------------------------------------------------------------ VGG19 glue: max-pool 2x2, relu mask
// NHWC f32 in -> planes (the next conv's operand) + optional f32.  Backward routes the gradient to the first
// maximum in (row, column) scan order of the window (ATen max_pool2d_with_indices: strict > keeps the first).
__global__ void __launch_bounds__(256)
    maxpool2x2_fwd_kernel(const float* __restrict__ x, float* __restrict__ yf, plane_t* __restrict__ yh,
                          plane_t* __restrict__ yl, int H, int W, int C, int cpad, int fmt, long total) {
  pdl_grid_sync();
  const int Ho = H / 2, Wo = W / 2;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    long p = e / C;
    const int ox = (int)(p % Wo);
    p /= Wo;
    const int oy = (int)(p % Ho);
    const long n = p / Ho;
    const float* s = x + (((long)n * H + 2 * oy) * W + 2 * ox) * C + c;
    float m = s[0];
    m = fmaxf(m, s[C]);
    m = fmaxf(m, s[(long)W * C]);
    m = fmaxf(m, s[(long)W * C + C]);
    const long op = ((long)n * Ho + oy) * Wo + ox;
    if (yf) yf[op * C + c] = m;
    if (yh) {
      plane_t h, l;
      split16(m, fmt, h, l);
      yh[op * cpad + c] = h;
      if (yl) yl[op * cpad + c] = l;
    }
  }
}

__global__ void __launch_bounds__(256)
    maxpool2x2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, int gy_cstride, float* __restrict__ gx,
                          int H, int W, int C, long total) {
  pdl_grid_sync();
  const int Ho = H / 2, Wo = W / 2;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    long p = e / C;
    const int ox = (int)(p % Wo);
    p /= Wo;
    const int oy = (int)(p % Ho);
    const long n = p / Ho;
    const long i00 = (((long)n * H + 2 * oy) * W + 2 * ox) * C + c;
    const long idx[4] = {i00, i00 + C, i00 + (long)W * C, i00 + (long)W * C + C};
    int best = 0;
    float m = x[idx[0]];
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const float v = x[idx[k]];
      if (v > m) { m = v; best = k; }
    }
    const float g = gy[(((long)n * Ho + oy) * Wo + ox) * gy_cstride + c];
#pragma unroll
    for (int k = 0; k < 4; ++k) gx[idx[k]] = k == best ? g : 0.f;
  }
}

// y[n,c,h,w] (+)= x[n,h,w,c] for a small channel count (gradient of an NCHW image produced by an NHWC kernel)
__global__ void __launch_bounds__(256)
    nhwc_to_nchw_add_kernel(const float* __restrict__ x, int cstride, float* __restrict__ y, int HW, int C, long total,
                            int accumulate) {
  pdl_grid_sync();
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int p = (int)(e % HW);
    const long r = e / HW;
    const int c = (int)(r % C);
    const long n = r / C;
    const float v = __ldg(x + (n * HW + p) * cstride + c);
    y[e] = accumulate ? y[e] + v : v;
  }
}

}  // namespace shineon

using namespace shineon;

extern "C" int shineon_nhwc_to_nchw_add(const float* x, int x_cstride, float* y, int N, int H, int W, int C, int accumulate,
                                        shineon_stream_t stream) {
  SHINEON_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && x_cstride >= C, "nhwc_to_nchw_add: bad arguments");
  const long total = (long)N * C * H * W;
  klaunch(nhwc_to_nchw_add_kernel, grid_x(total, 256), 256, 0, (cudaStream_t)stream, x, x_cstride, y, H * W, C, total, accumulate);
  return after_launch("nhwc_to_nchw_add_kernel");
}

extern "C" int shineon_instnorm_act_bwd(const float* x, const double* stats_fwd, const float* g1, const float* g2,
                                        float* gx_f32, void* gx_hi, void* gx_lo, double* stats_ws, int N, int H, int W,
                                        int C, int cpad, float eps, int do_norm, int act, float act_param, int plane_fmt,
                                        shineon_stream_t stream_) {
  SHINEON_REQUIRE(x && g1 && (gx_f32 || gx_hi), "instnorm_act_bwd: null pointer");
  SHINEON_REQUIRE(plane_fmt == SHINEON_FMT_BF16 || plane_fmt == SHINEON_FMT_FP16, "instnorm_act_bwd: plane_fmt %d", plane_fmt);
  SHINEON_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && C > 0, "instnorm_act_bwd: bad shape");
  SHINEON_REQUIRE(!do_norm || (stats_fwd && stats_ws), "instnorm_act_bwd: statistics buffers required");
  SHINEON_REQUIRE(!gx_hi || cpad >= C, "instnorm_act_bwd: cpad < C");
  SHINEON_REQUIRE(4 * C * sizeof(float) <= 48 * 1024, "instnorm_act_bwd: C too large");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int HW = H * W;
  if (do_norm) {
    cudaError_t e = cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * (size_t)N * C, stream);
    if (e != cudaSuccess) return fail(SHINEON_ERR_CUDA, "instnorm_act_bwd memset: %s", cudaGetErrorString(e));
    const int pix_per_cta = 128;
    dim3 grid(cdiv(HW, pix_per_cta), cdiv(C, 32), N);
    klaunch(instnorm_bwd_stats_kernel, grid, 256, 0, stream, x, stats_fwd, g1, g2, stats_ws, HW, C, pix_per_cta, eps, act, act_param);
    int rc = after_launch("instnorm_bwd_stats_kernel");
    if (rc) return rc;
  }
  const int vec = (C % 4 == 0) ? 4 : 1;
  SHINEON_REQUIRE((long)HW * (C / vec) < (1l << 31), "instnorm_act_bwd: image too large");
  dim3 grid(grid_x((long)HW * (C / vec), 256), N);
  const size_t sm = 4 * C * sizeof(float);
  if (vec == 4)
    klaunch(instnorm_bwd_apply_kernel<4>, grid, 256, sm, stream, x, stats_fwd, stats_ws, g1, g2, gx_f32, (plane_t*)gx_hi, (plane_t*)gx_lo,
                                                            HW, C, cpad, eps, do_norm, act, act_param, plane_fmt);
  else
    klaunch(instnorm_bwd_apply_kernel<1>, grid, 256, sm, stream, x, stats_fwd, stats_ws, g1, g2, gx_f32, (plane_t*)gx_hi, (plane_t*)gx_lo,
                                                            HW, C, cpad, eps, do_norm, act, act_param, plane_fmt);
  return after_launch("instnorm_bwd_apply_kernel");
}

extern "C" int shineon_act_bwd(const float* z, const float* g1, const float* g2, float* gz, long n, int act, float act_param,
                               shineon_stream_t stream) {
  SHINEON_REQUIRE(z && g1 && gz && n > 0, "act_bwd: bad arguments");
  klaunch(act_bwd_kernel, grid_x(n, 256), 256, 0, (cudaStream_t)stream, z, g1, g2, gz, n, act, act_param);
  return after_launch("act_bwd_kernel");
}

extern "C" int shineon_upsample2x_cat_bwd(const float* g_up, int g_cstride, float* g0, int C0, float* g1, int C1, int N, int H,
                                          int W, shineon_stream_t stream) {
  SHINEON_REQUIRE(g_up && g0 && C0 > 0 && (g1 != nullptr) == (C1 > 0), "upsample2x_cat_bwd: bad arguments");
  SHINEON_REQUIRE(g_cstride >= C0 + C1 && N > 0 && N <= 65535 && H > 0 && H <= 65535 && W > 0, "upsample2x_cat_bwd: bad shape");
  dim3 grid(cdiv(W * ((C0 + C1 + 3) / 4), 256), H, N);
  klaunch(upsample2x_cat_bwd_kernel, grid, 256, 0, (cudaStream_t)stream, g_up, g_cstride, g0, C0, g1, C1, H, W);
  return after_launch("upsample2x_cat_bwd_kernel");
}

extern "C" size_t shineon_sagan_attention_bwd_workspace_bytes(int N, int HW) {
  return (size_t)2 * N * HW * HW * sizeof(float) + 64;
}

extern "C" int shineon_sagan_attention_bwd(const float* qkv, const float* gamma, const float* g_out, float* g_qkv,
                                           float* g_gamma, void* workspace, size_t workspace_bytes, int N, int HW, int C,
                                           int Cq, float beta_gamma, shineon_stream_t stream_) {
  SHINEON_REQUIRE(qkv && gamma && g_out && g_qkv && g_gamma && workspace, "sagan_attention_bwd: null pointer");
  SHINEON_REQUIRE(N > 0 && N <= 65535 && HW > 0 && C > 0 && Cq > 0, "sagan_attention_bwd: bad shape");
  SHINEON_REQUIRE(workspace_bytes >= shineon_sagan_attention_bwd_workspace_bytes(N, HW), "sagan_attention_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream_;
  double* acc = (double*)workspace;
  float* Abuf = (float*)((char*)workspace + 64);
  float* dEbuf = Abuf + (size_t)N * HW * HW;
  cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(double), st);
  if (e != cudaSuccess) return fail(SHINEON_ERR_CUDA, "sagan_attention_bwd memset: %s", cudaGetErrorString(e));
  constexpr int QT = 8;
  const size_t smA = sizeof(float) * ((size_t)QT * Cq + 2 * (size_t)QT * HW + QT * 4 + QT);
  SHINEON_REQUIRE(smA <= 48 * 1024 && 2 * HW * sizeof(float) <= 48 * 1024, "sagan_attention_bwd: HW=%d too large for this kernel", HW);
  klaunch(sagan_bwd_a_kernel<QT>, dim3(cdiv(HW, QT), N), 128, smA, st, qkv, gamma, g_out, Abuf, dEbuf, g_qkv, acc, HW, C, Cq);
  int rc = after_launch("sagan_bwd_a_kernel");
  if (rc) return rc;
  klaunch(sagan_bwd_b_kernel, dim3(HW, N), 128, 2 * HW * sizeof(float), st, qkv, gamma, g_out, Abuf, dEbuf, g_qkv, HW, C, Cq);
  rc = after_launch("sagan_bwd_b_kernel");
  if (rc) return rc;
  klaunch(scalar_finish_kernel, 1, 32, 0, st, acc, g_gamma, 1.f, beta_gamma);
  return after_launch("scalar_finish_kernel");
}

extern "C" int shineon_tom_compose_bwd(const float* unet_out, int Cout, const float* cloth, const float* warped_prev,
                                       const float* g_rendereds, const float* g_masks, const float* g_tryons,
                                       const float* g_flow_masks, float* g_unet_out, float* g_warped_prev, int B, int H,
                                       int W, int n_frames, int frame, int flow_warp, shineon_stream_t stream) {
  SHINEON_REQUIRE(unet_out && cloth && g_unet_out, "tom_compose_bwd: null pointer");
  SHINEON_REQUIRE(n_frames >= 1 && frame >= 0 && frame < n_frames, "tom_compose_bwd: frame %d of %d", frame, n_frames);
  SHINEON_REQUIRE(Cout == (flow_warp ? 5 : 4) * n_frames, "tom_compose_bwd: Cout %d", Cout);
  SHINEON_REQUIRE(!warped_prev || flow_warp, "tom_compose_bwd: warped_prev needs flow_warp");
  SHINEON_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, "tom_compose_bwd: bad shape");
  dim3 grid(grid_x((long)H * W, 256), B);
  klaunch(tom_compose_bwd_kernel, grid, 256, 0, (cudaStream_t)stream, unet_out, Cout, cloth, warped_prev, g_rendereds, g_masks,
                                                                g_tryons, g_flow_masks, g_unet_out, g_warped_prev, H * W,
                                                                n_frames, frame, flow_warp);
  return after_launch("tom_compose_bwd_kernel");
}

extern "C" int shineon_l1_loss(const float* a, const float* b, float* grad_a, float* loss, void* workspace, long n,
                               float loss_weight, float beta_loss, int accumulate_grad, shineon_stream_t stream_) {
  SHINEON_REQUIRE(a && b && loss && workspace && n > 0, "l1_loss: bad arguments");
  cudaStream_t st = (cudaStream_t)stream_;
  cudaError_t e = cudaMemsetAsync(workspace, 0, sizeof(double), st);
  if (e != cudaSuccess) return fail(SHINEON_ERR_CUDA, "l1_loss memset: %s", cudaGetErrorString(e));
  int blocks = grid_x(n, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  klaunch(l1_loss_kernel, blocks, 256, 0, st, a, b, grad_a, (double*)workspace, n, loss_weight / (float)n, accumulate_grad);
  int rc = after_launch("l1_loss_kernel");
  if (rc) return rc;
  klaunch(scalar_finish_kernel, 1, 32, 0, st, (const double*)workspace, loss, loss_weight / (float)n, beta_loss);
  return after_launch("scalar_finish_kernel");
}

extern "C" int shineon_maxpool2x2_fwd(const float* x, float* y_f32, void* y_hi, void* y_lo, int N, int H, int W, int C,
                                      int cpad, int plane_fmt, shineon_stream_t stream) {
  SHINEON_REQUIRE(x && (y_f32 || y_hi) && N > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0, "maxpool2x2_fwd: bad arguments");
  SHINEON_REQUIRE(!y_hi || cpad >= C, "maxpool2x2_fwd: cpad < C");
  const long total = (long)N * (H / 2) * (W / 2) * C;
  klaunch(maxpool2x2_fwd_kernel, grid_x(total, 256), 256, 0, (cudaStream_t)stream, x, y_f32, (plane_t*)y_hi, (plane_t*)y_lo, H, W, C,
                                                                             cpad, plane_fmt, total);
  return after_launch("maxpool2x2_fwd_kernel");
}

extern "C" int shineon_maxpool2x2_bwd(const float* x, const float* g_y, int g_cstride, float* g_x, int N, int H, int W, int C,
                                      shineon_stream_t stream) {
  SHINEON_REQUIRE(x && g_y && g_x && N > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0 && g_cstride >= C, "maxpool2x2_bwd: bad arguments");
  const long total = (long)N * (H / 2) * (W / 2) * C;
  klaunch(maxpool2x2_bwd_kernel, grid_x(total, 256), 256, 0, (cudaStream_t)stream, x, g_y, g_cstride, g_x, H, W, C, total);
  return after_launch("maxpool2x2_bwd_kernel");
}
