// SAGAN self-attention core (reference: models/networks/attention/sagan.py:29-53).
// The three 1x1 projections are one tcgen05 conv (Cout = 2*Cq + C); this kernel does
//   energy[i,j] = <q_i, k_j>;  A = softmax_j(energy);  o_i[c] = sum_j A[i,j] v_j[c];  y = act(gamma*o + x)
// for one query pixel i per CTA.  N = H*W <= 192 on the ShineOn U-Net (bottom two levels), so the whole
// key/value set of an image stays L1/L2 resident; fp32 throughout.
#include "common.cuh"

namespace shineon {

__global__ void __launch_bounds__(128)
    sagan_attention_kernel(const float* __restrict__ qkv, const float* __restrict__ x, const float* __restrict__ gamma,
                           float* __restrict__ yf, plane_t* __restrict__ yh, plane_t* __restrict__ yl,
                           int HW, int C, int Cq, int cpad, int act, float act_param, int fmt) {
  extern __shared__ float sm[];  // q[Cq] | e[HW] | red[32]
  float* sq = sm;
  float* se = sm + Cq;
  float* red = se + HW;
  const int n = blockIdx.y, i = blockIdx.x;
  const int ld = 2 * Cq + C;
  const float* base = qkv + (long)n * HW * ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int c = tid; c < Cq; c += blockDim.x) sq[c] = base[(long)i * ld + c];
  __syncthreads();

  // energies
  float lmax = -INFINITY;
  for (int j = tid; j < HW; j += blockDim.x) {
    const float* kj = base + (long)j * ld + Cq;
    float acc = 0.f;
    if ((Cq & 3) == 0) {
      for (int c = 0; c < Cq; c += 4) {
        float4 k4 = *reinterpret_cast<const float4*>(kj + c);
        acc = fmaf(sq[c], k4.x, acc);
        acc = fmaf(sq[c + 1], k4.y, acc);
        acc = fmaf(sq[c + 2], k4.z, acc);
        acc = fmaf(sq[c + 3], k4.w, acc);
      }
    } else {
      for (int c = 0; c < Cq; ++c) acc = fmaf(sq[c], kj[c], acc);
    }
    se[j] = acc;
    lmax = fmaxf(lmax, acc);
  }
  lmax = warp_max(lmax);
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  float gmax = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) gmax = fmaxf(gmax, red[w]);
  __syncthreads();
  float lsum = 0.f;
  for (int j = tid; j < HW; j += blockDim.x) {
    float e = expf(se[j] - gmax);
    se[j] = e;
    lsum += e;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) red[warp] = lsum;
  __syncthreads();
  float gsum = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) gsum += red[w];
  const float inv = 1.f / gsum;
  const float g = __ldg(gamma);

  // o[c] = sum_j A[j] * v_j[c]; threads stride channels (coalesced rows of V)
  const float* vbase = base + 2 * Cq;
  for (int c = tid; c < C; c += blockDim.x) {
    float a0 = 0.f, a1 = 0.f;
    int j = 0;
    for (; j + 1 < HW; j += 2) {
      a0 = fmaf(se[j], vbase[(long)j * ld + c], a0);
      a1 = fmaf(se[j + 1], vbase[(long)(j + 1) * ld + c], a1);
    }
    if (j < HW) a0 = fmaf(se[j], vbase[(long)j * ld + c], a0);
    const float o = (a0 + a1) * inv;
    const long xi = ((long)n * HW + i) * C + c;
    const float v = apply_act(g * o + x[xi], act, act_param);  // sagan.py:53
    if (yf) yf[xi] = v;
    if (yh) {
      plane_t h, l;
      split16(v, fmt, h, l);
      const long po = ((long)n * HW + i) * cpad + c;
      yh[po] = h;
      if (yl) yl[po] = l;
    }
  }
}

}  // namespace shineon

using namespace shineon;

extern "C" int shineon_sagan_attention(const float* qkv, const float* x, const float* gamma, float* y_f32, void* y_hi,
                                       void* y_lo, int N, int HW, int C, int Cq, int cpad, int act, float act_param,
                                       int plane_fmt, shineon_stream_t stream) {
  SHINEON_REQUIRE(plane_fmt == SHINEON_FMT_BF16 || plane_fmt == SHINEON_FMT_FP16, "sagan_attention: plane_fmt %d", plane_fmt);
  SHINEON_REQUIRE(qkv && x && gamma && (y_f32 || y_hi), "sagan_attention: null pointer");
  SHINEON_REQUIRE(N > 0 && N <= 65535 && HW > 0 && C > 0 && Cq > 0, "sagan_attention: bad shape");
  SHINEON_REQUIRE(!y_hi || cpad >= C, "sagan_attention: cpad < C");
  size_t smem = sizeof(float) * (size_t)(Cq + HW + 32);
  if (smem > 48 * 1024) return fail(SHINEON_ERR_UNSUPPORTED, "sagan_attention: HW=%d too large for this kernel", HW);
  SHINEON_REQUIRE(((2 * Cq + C) & 3) == 0 || (Cq & 3) != 0, "sagan_attention: row stride must keep float4 alignment");
  dim3 grid(HW, N);
  sagan_attention_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(qkv, x, gamma, y_f32, (plane_t*)y_hi,
                                                                   (plane_t*)y_lo, HW, C, Cq, cpad, act, act_param, plane_fmt);
  return after_launch("sagan_attention_kernel");
}
