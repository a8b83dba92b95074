// SAGAN self-attention core (reference: models/networks/attention/sagan.py:29-53).
// The three 1x1 projections are one tcgen05 conv (Cout = 2*Cq + C); this kernel does
//   energy[i,j] = <q_i, k_j>;  A = softmax_j(energy);  o_i[c] = sum_j A[i,j] v_j[c];  y = act(gamma*o + x)
// for one query pixel i per CTA.  N = H*W <= 192 on the ShineOn U-Net (bottom two levels), so the whole
// key/value set of an image stays L1/L2 resident; fp32 throughout.
#include "common.cuh"

namespace shineon {

// QT query pixels per CTA: K and V rows are read once per CTA and reused for all QT queries (the first version
// used one query per CTA and was L2-bandwidth bound re-reading V: 15k CTAs x 440 KB).
template <int QT>
__global__ void __launch_bounds__(128)
    sagan_attention_kernel(const float* __restrict__ qkv, const float* __restrict__ x, const float* __restrict__ gamma,
                           float* __restrict__ yf, plane_t* __restrict__ yh, plane_t* __restrict__ yl,
                           int HW, int C, int Cq, int cpad, int act, float act_param, int fmt) {
  extern __shared__ float sm[];  // q[QT][Cq] | e[QT][HW] | inv[QT]
  float* sq = sm;
  float* se = sm + QT * Cq;
  float* sinv = se + QT * HW;
  const int n = blockIdx.y, i0 = blockIdx.x * QT;
  const int nq = min(QT, HW - i0);
  const int ld = 2 * Cq + C;
  const float* base = qkv + (long)n * HW * ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int e = tid; e < QT * Cq; e += 128) {
    const int q = e / Cq, c = e - q * Cq;
    sq[e] = q < nq ? base[(long)(i0 + q) * ld + c] : 0.f;
  }
  __syncthreads();

  // energies e[q][j] = <q_q, k_j>
  for (int j = tid; j < HW; j += 128) {
    const float* kj = base + (long)j * ld + Cq;
    float acc[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) acc[q] = 0.f;
    if ((Cq & 3) == 0) {
      for (int c = 0; c < Cq; c += 4) {
        const float4 k4 = *reinterpret_cast<const float4*>(kj + c);
#pragma unroll
        for (int q = 0; q < QT; ++q) {
          const float* qq = sq + q * Cq + c;
          acc[q] = fmaf(qq[0], k4.x, acc[q]);
          acc[q] = fmaf(qq[1], k4.y, acc[q]);
          acc[q] = fmaf(qq[2], k4.z, acc[q]);
          acc[q] = fmaf(qq[3], k4.w, acc[q]);
        }
      }
    } else {
      for (int c = 0; c < Cq; ++c) {
        const float kv = kj[c];
#pragma unroll
        for (int q = 0; q < QT; ++q) acc[q] = fmaf(sq[q * Cq + c], kv, acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < QT; ++q) se[q * HW + j] = acc[q];
  }
  __syncthreads();

  // softmax over j, one warp per query
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
    if (lane == 0) sinv[q] = 1.f / sum;
  }
  __syncthreads();

  // o[q][c] = sum_j A[q][j] v_j[c]; 4 consecutive channels per thread (coalesced float4 rows of V)
  const float g = __ldg(gamma);
  const float* vbase = base + 2 * Cq;
  const bool vec = ((C & 3) == 0) && ((ld & 3) == 0) && (((2 * Cq) & 3) == 0);
  const int cstep = vec ? 4 : 1;
  for (int c = tid * cstep; c < C; c += 128 * cstep) {
    float acc[QT][4];
#pragma unroll
    for (int q = 0; q < QT; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;
    for (int j = 0; j < HW; ++j) {
      float4 v4;
      if (vec)
        v4 = *reinterpret_cast<const float4*>(vbase + (long)j * ld + c);
      else
        v4 = make_float4(vbase[(long)j * ld + c], 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < QT; ++q) {
        const float aw = se[q * HW + j];
        acc[q][0] = fmaf(aw, v4.x, acc[q][0]);
        acc[q][1] = fmaf(aw, v4.y, acc[q][1]);
        acc[q][2] = fmaf(aw, v4.z, acc[q][2]);
        acc[q][3] = fmaf(aw, v4.w, acc[q][3]);
      }
    }
#pragma unroll
    for (int q = 0; q < QT; ++q) {
      if (q >= nq) break;
      const long pix = (long)n * HW + i0 + q;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k >= cstep) break;
        const float o = acc[q][k] * sinv[q];
        const float v = apply_act(g * o + x[pix * C + c + k], act, act_param);  // sagan.py:53
        if (yf) yf[pix * C + c + k] = v;
        if (yh) {
          plane_t h, l;
          split16(v, fmt, h, l);
          yh[pix * cpad + c + k] = h;
          if (yl) yl[pix * cpad + c + k] = l;
        }
      }
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
  SHINEON_REQUIRE(((2 * Cq + C) & 3) == 0 || (Cq & 3) != 0, "sagan_attention: row stride must keep float4 alignment");
  auto smem_for = [&](int qt) { return sizeof(float) * ((size_t)qt * Cq + (size_t)qt * HW + qt); };
  cudaStream_t st = (cudaStream_t)stream;
#define SHINEON_ATT(QT_)                                                                                            \
  sagan_attention_kernel<QT_><<<dim3(cdiv(HW, QT_), N), 128, smem_for(QT_), st>>>(                                   \
      qkv, x, gamma, y_f32, (plane_t*)y_hi, (plane_t*)y_lo, HW, C, Cq, cpad, act, act_param, plane_fmt)
  if (smem_for(16) <= 40 * 1024) SHINEON_ATT(16);
  else if (smem_for(4) <= 40 * 1024) SHINEON_ATT(4);
  else if (smem_for(1) <= 48 * 1024) SHINEON_ATT(1);
  else return fail(SHINEON_ERR_UNSUPPORTED, "sagan_attention: HW=%d too large for this kernel", HW);
#undef SHINEON_ATT
  return after_launch("sagan_attention_kernel");
}
