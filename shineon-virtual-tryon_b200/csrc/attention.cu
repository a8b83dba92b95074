// SAGAN self-attention core (reference: models/networks/attention/sagan.py:29-53).
// The three 1x1 projections are one tcgen05 conv (Cout = 2*Cq + C); this kernel does
//   energy[i,j] = <q_i, k_j>;  A = softmax_j(energy);  o_i[c] = sum_j A[i,j] v_j[c];  y = act(gamma*o + x)
// for one query pixel i per CTA.  N = H*W <= 192 on the ShineOn U-Net (bottom two levels), so the whole
// key/value set of an image stays L1/L2 resident; fp32 throughout.
#include <stdlib.h>

#include "tcgen05.cuh"

namespace shineon {

// QT query pixels per CTA: K and V rows are read once per CTA and reused for all QT queries (the first version
// used one query per CTA and was L2-bandwidth bound re-reading V: 15k CTAs x 440 KB).
template <int QT>
__global__ void __launch_bounds__(128)
    sagan_attention_kernel(const float* __restrict__ qkv, const float* __restrict__ x, const float* __restrict__ gamma,
                           float* __restrict__ yf, plane_t* __restrict__ yh, plane_t* __restrict__ yl,
                           int HW, int C, int Cq, int cpad, int act, float act_param, int fmt) {
  pdl_grid_sync();
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

// One query row x 8 channels of the epilogue: y = act(gamma * o / sum + x) -> f32 and / or 16-bit planes.  Deliberately
// NOT inlined: with the runtime activation switch inlined for all 64 accumulators the tiled kernel was 32k SASS
// instructions, most of them executed once, and a fifth of its stall samples were instruction-cache misses
// (stall_no_inst, profiles/r01_attention.md).
__device__ __noinline__ void attn_store_row8(float4 a0, float4 a1, float inv, float g, const float* __restrict__ xrow,
                                             float* yf_row, plane_t* yh_row, plane_t* yl_row, int act, float act_param, int fmt) {
  const float4 xa = __ldg(reinterpret_cast<const float4*>(xrow));
  const float4 xb = __ldg(reinterpret_cast<const float4*>(xrow + 4));
  const float o[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  const float xr[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = g * (o[k] * inv) + xr[k];  // sagan.py:53
  switch (act) {  // uniform
    case SHINEON_ACT_NONE: break;
    case SHINEON_ACT_GELU:
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = apply_act(v[k], SHINEON_ACT_GELU, act_param);
      break;
    case SHINEON_ACT_RELU:
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = apply_act(v[k], SHINEON_ACT_RELU, act_param);
      break;
    default:
#pragma unroll 1
      for (int k = 0; k < 8; ++k) v[k] = apply_act(v[k], act, act_param);
      break;
  }
  if (yf_row) {
    reinterpret_cast<float4*>(yf_row)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(yf_row)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (yh_row) {
    plane_t h[8], l[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) split16(v[k], fmt, h[k], l[k]);
    *reinterpret_cast<uint4*>(yh_row) =
        make_uint4((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16),
                   (uint32_t)h[4] | ((uint32_t)h[5] << 16), (uint32_t)h[6] | ((uint32_t)h[7] << 16));
    if (yl_row)
      *reinterpret_cast<uint4*>(yl_row) =
          make_uint4((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16),
                     (uint32_t)l[4] | ((uint32_t)l[5] << 16), (uint32_t)l[6] | ((uint32_t)l[7] << 16));
  }
}

// Register-tiled version (the default when the row strides keep 16-byte alignment): 256 threads, QB queries per CTA.
//   energies  : one (key j, 8-query group) item per thread, q rows broadcast from shared memory;
//   softmax   : thread = (query, 1/8 of the keys), partial max / sum combined through shared memory; probabilities are
//               kept TRANSPOSED [HW][QB] so the next phase reads 8 queries with two 128-bit broadcast loads;
//   P.V       : each thread owns 8 queries x 8 channels (64 accumulators): per key 2 LDG.128 of V (coalesced 2 KB rows,
//               shared by the CTA's query groups through L1) + 2 LDS.128 of P feed 64 FMAs.
// The first kernel below (16 queries x 4 channels per thread, 128 threads) spent 16 scalar LDS per 64 FMAs and ran
// latency-bound at ~15 % of the fp32 FMA rate (profiles/r01_launches_summary.md: 0.30 ms for N = 192, C = 512, 80 images).
template <int QB>
__global__ void __launch_bounds__(256)
    sagan_attention_tiled_kernel(const float* __restrict__ qkv, const float* __restrict__ x, const float* __restrict__ gamma,
                                 float* __restrict__ yf, plane_t* __restrict__ yh, plane_t* __restrict__ yl,
                                 int HW, int C, int Cq, int cpad, int act, float act_param, int fmt, int kc) {
  pdl_grid_sync();
  constexpr int QG = QB / 8;       // 8-query groups
  constexpr int PARTS = 256 / QB;  // key partitions of the softmax phase
  extern __shared__ __align__(16) float sm[];
  float* sq = sm;                    // [QB][Cq]
  float* sp = sm + QB * Cq;          // [HW][QB] energies -> un-normalised probabilities
  float* sred = sp + (size_t)HW * QB;  // [PARTS][QB] partial max / sum
  float* sinv = sred + PARTS * QB;   // [QB]
  const int n = blockIdx.y, i0 = blockIdx.x * QB;
  const int nq = min(QB, HW - i0);
  const int ld = 2 * Cq + C;
  const float* base = qkv + (long)n * HW * ld;
  const int tid = threadIdx.x;

  for (int e = tid; e < QB * Cq; e += 256) {
    const int q = e / Cq, c = e - q * Cq;
    sq[e] = q < nq ? __ldg(base + (long)(i0 + q) * ld + c) : 0.f;
  }
  __syncthreads();

  for (int item = tid; item < HW * QG; item += 256) {
    const int qg = item / HW, j = item - qg * HW;
    const float* kj = base + (long)j * ld + Cq;
    const float* q0 = sq + qg * 8 * Cq;
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    for (int c = 0; c < Cq; c += 4) {
      const float4 k4 = __ldg(reinterpret_cast<const float4*>(kj + c));
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 q4 = *reinterpret_cast<const float4*>(q0 + q * Cq + c);
        acc[q] = fmaf(q4.x, k4.x, acc[q]);
        acc[q] = fmaf(q4.y, k4.y, acc[q]);
        acc[q] = fmaf(q4.z, k4.z, acc[q]);
        acc[q] = fmaf(q4.w, k4.w, acc[q]);
      }
    }
    float4* dst = reinterpret_cast<float4*>(sp + (size_t)j * QB + qg * 8);
    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  __syncthreads();

  {  // softmax over keys; thread = (query q, key partition part)
    const int q = tid % QB, part = tid / QB;
    float m = -INFINITY;
    for (int j = part; j < HW; j += PARTS) m = fmaxf(m, sp[(size_t)j * QB + q]);
    sred[part * QB + q] = m;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PARTS; ++k) m = fmaxf(m, sred[k * QB + q]);
    __syncthreads();
    float sum = 0.f;
    for (int j = part; j < HW; j += PARTS) {
      const float pr = expf(sp[(size_t)j * QB + q] - m);
      sp[(size_t)j * QB + q] = pr;
      sum += pr;
    }
    sred[part * QB + q] = sum;
    __syncthreads();
    if (part == 0) {
      float tot = 0.f;
#pragma unroll
      for (int k = 0; k < PARTS; ++k) tot += sred[k * QB + q];
      sinv[q] = 1.f / tot;
    }
    __syncthreads();
  }

  const float g = __ldg(gamma);
  const float* vbase = base + 2 * Cq;
  constexpr int CGP = 256 / QG;  // channel groups (of 8) per pass
  const int qg = tid / CGP;
  const int q_lo = qg * 8;
  if (kc > 0) {
    // ---- P.V with the value rows staged through shared memory: thread 0 streams chunks of kc keys (each row one
    // cp.async.bulk of C floats, two chunks in flight) while all threads accumulate from the other buffer.  Reading V
    // straight from global (the loop below) left only ~4 loads in flight per thread: latency-bound at 20 % of the FMA rate.
    float* sv = sinv + QB;  // [2][kc][C]; 16-byte aligned (every section above is a multiple of 4 floats)
    __shared__ __align__(8) uint64_t vbar[2];
    const uint32_t bar0 = smem_u32(&vbar[0]);
    if (tid == 0) {
      mbar_init(bar0, 1);
      mbar_init(bar0 + 8, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nchunks = (HW + kc - 1) / kc;
    auto issue = [&](int ch) {
      const int buf = ch & 1, j0 = ch * kc;
      const int nk = min(kc, HW - j0);
      const uint32_t bar = bar0 + 8 * buf;
      mbar_expect_tx(bar, (uint32_t)nk * C * 4);
      for (int t = 0; t < nk; ++t)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sv + ((size_t)buf * kc + t) * C)), "l"(vbase + (long)(j0 + t) * ld), "r"(C * 4), "r"(bar)
                     : "memory");
    };
    if (tid == 0) {
      issue(0);
      if (nchunks > 1) issue(1);
    }
    const int c0 = (tid % CGP) * 8;
    const bool active = c0 < C;
    float acc[8][8];
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[q][k] = 0.f;
    for (int ch = 0; ch < nchunks; ++ch) {
      const int buf = ch & 1, j0 = ch * kc;
      const int nk = min(kc, HW - j0);
      mbar_wait(bar0 + 8 * buf, (ch >> 1) & 1);
      if (active) {
        const float* vp = sv + (size_t)buf * kc * C + c0;
        const float* pp = sp + (size_t)j0 * QB + q_lo;
#pragma unroll 2
        for (int t = 0; t < nk; ++t) {
          const float4 va = *reinterpret_cast<const float4*>(vp + (size_t)t * C);
          const float4 vb = *reinterpret_cast<const float4*>(vp + (size_t)t * C + 4);
          const float4 pa = *reinterpret_cast<const float4*>(pp + (size_t)t * QB);
          const float4 pb = *reinterpret_cast<const float4*>(pp + (size_t)t * QB + 4);
          const float v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
          const float pq[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
          for (int q = 0; q < 8; ++q)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[q][k] = fmaf(pq[q], v[k], acc[q][k]);
        }
      }
      __syncthreads();  // everyone is done with this buffer
      if (tid == 0 && ch + 2 < nchunks) issue(ch + 2);
    }
    if (active) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q_lo + q >= nq) break;
        const long pix = (long)n * HW + i0 + q_lo + q;
        attn_store_row8(make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]), make_float4(acc[q][4], acc[q][5], acc[q][6], acc[q][7]),
                        sinv[q_lo + q], g, x + pix * C + c0, yf ? yf + pix * C + c0 : nullptr, yh ? yh + pix * cpad + c0 : nullptr,
                        yl ? yl + pix * cpad + c0 : nullptr, act, act_param, fmt);
      }
    }
    return;
  }
  for (int c0 = (tid % CGP) * 8; c0 < C; c0 += CGP * 8) {
    float acc[8][8];
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[q][k] = 0.f;
    const float* vp = vbase + c0;
    const float* pp = sp + q_lo;
#pragma unroll 2
    for (int j = 0; j < HW; ++j) {
      const float4 va = __ldg(reinterpret_cast<const float4*>(vp + (long)j * ld));
      const float4 vb = __ldg(reinterpret_cast<const float4*>(vp + (long)j * ld + 4));
      const float4 pa = *reinterpret_cast<const float4*>(pp + (size_t)j * QB);
      const float4 pb = *reinterpret_cast<const float4*>(pp + (size_t)j * QB + 4);
      const float v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
      const float pq[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[q][k] = fmaf(pq[q], v[k], acc[q][k]);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (q_lo + q >= nq) break;
      const long pix = (long)n * HW + i0 + q_lo + q;
      attn_store_row8(make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]), make_float4(acc[q][4], acc[q][5], acc[q][6], acc[q][7]),
                      sinv[q_lo + q], g, x + pix * C + c0, yf ? yf + pix * C + c0 : nullptr, yh ? yh + pix * cpad + c0 : nullptr,
                      yl ? yl + pix * cpad + c0 : nullptr, act, act_param, fmt);
    }
  }
}

}  // namespace shineon

using namespace shineon;

static inline int ld_floats(int C, int Cq) { return 2 * Cq + C; }

int shineon_sagan_attention_tc(const float* qkv, const float* x, const float* gamma, float* y_f32, void* y_hi, void* y_lo,
                               int N, int HW, int C, int Cq, int cpad, int act, float act_param, int plane_fmt,
                               cudaStream_t stream);  // attention_tc.cu

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
  // tensor-core kernel (attention_tc.cu) for the U-Net's shapes; > 0 = shape does not fit.  SHINEON_ATTENTION_TC=0 keeps the
  // CUDA-core kernels below (A/B measurements, the fp32 cross-check in the tests)
  static const bool use_tc = [] { const char* e = getenv("SHINEON_ATTENTION_TC"); return !(e && e[0] == '0'); }();
  if (use_tc) {
    const int rc = shineon_sagan_attention_tc(qkv, x, gamma, y_f32, y_hi, y_lo, N, HW, C, Cq, cpad, act, act_param, plane_fmt, st);
    if (rc <= 0) return rc;
  }
  // register-tiled kernel: needs 16-byte aligned q/k/v/x rows and plane rows (8 channels = one 128-bit store)
  const bool aligned = (C % 8 == 0) && (Cq % 4 == 0) && (!y_hi || cpad % 8 == 0) &&
                       ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y_f32) |
                         reinterpret_cast<uintptr_t>(y_hi) | reinterpret_cast<uintptr_t>(y_lo)) % 16 == 0);
  if (aligned) {
    constexpr int QB = 32;
    size_t smem = sizeof(float) * ((size_t)QB * Cq + (size_t)HW * QB + (256 / QB) * QB + QB);
    // value rows staged through shared memory (two buffers of kc keys) when one channel pass covers C and it fits
    int kc = 0;
    if (C <= (256 / (QB / 8)) * 8 && (ld_floats(C, Cq) % 4) == 0) {
      kc = 16;
      while (kc > 2 && smem + 2 * (size_t)kc * C * sizeof(float) > 100 * 1024) kc /= 2;
      if (smem + 2 * (size_t)kc * C * sizeof(float) > 100 * 1024) kc = 0;
      if (kc > HW) kc = HW;
    }
    smem += 2 * (size_t)kc * C * sizeof(float);
    static size_t opted = 48 * 1024;  // dynamic shared memory this kernel has been opted into
    if (smem <= 200 * 1024) {
      if (smem > opted) {
        cudaError_t e = cudaFuncSetAttribute(sagan_attention_tiled_kernel<QB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(SHINEON_ERR_CUDA, "sagan_attention: shared memory opt-in: %s", cudaGetErrorString(e));
        opted = smem;
      }
      klaunch(sagan_attention_tiled_kernel<QB>, dim3(cdiv(HW, QB), N), 256, smem, st, 
          qkv, x, gamma, y_f32, (plane_t*)y_hi, (plane_t*)y_lo, HW, C, Cq, cpad, act, act_param, plane_fmt, kc);
      return after_launch("sagan_attention_tiled_kernel");
    }
  }
#define SHINEON_ATT(QT_)                                                                                            \
  klaunch(sagan_attention_kernel<QT_>, dim3(cdiv(HW, QT_), N), 128, smem_for(QT_), st,                                    \
      qkv, x, gamma, y_f32, (plane_t*)y_hi, (plane_t*)y_lo, HW, C, Cq, cpad, act, act_param, plane_fmt)
  if (smem_for(16) <= 40 * 1024) SHINEON_ATT(16);
  else if (smem_for(4) <= 40 * 1024) SHINEON_ATT(4);
  else if (smem_for(1) <= 48 * 1024) SHINEON_ATT(1);
  else return fail(SHINEON_ERR_UNSUPPORTED, "sagan_attention: HW=%d too large for this kernel", HW);
#undef SHINEON_ATT
  return after_launch("sagan_attention_kernel");
}
