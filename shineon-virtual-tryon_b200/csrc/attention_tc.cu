// SAGAN self-attention core on the 5th-gen tensor cores (reference: models/networks/attention/sagan.py:29-53).
//
//   energy = q^T k   [HW x HW, K = Cq]      A = softmax_j(energy)      o = A v   [HW x C, K = HW]      y = act(gamma o + x)
//
// One CTA per (image, 128-query tile).  Both contractions are tcgen05.mma with fp16 hi/lo split operands (three MMAs per
// product: hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM -- the same fp32-grade scheme as the convolutions, DESIGN.md
// section 4); the operands are produced IN the kernel from the f32 q|k|v projection (the output of the fused 1x1 conv):
//   * Q [128 x Cq] and K [HWp x Cq]: K-major SWIZZLE_128B tiles (row = pixel, 128-byte row = 64 channels);
//   * energies land in TMEM columns [0, HWp); the softmax warps (thread = query row = TMEM lane) take the row maximum,
//     exponentiate and write the un-normalised probabilities as the K-major A operand P [128 x HWp] (K = key index);
//   * V is consumed as an MN-MAJOR B operand: the smem tile is V as it lies in memory, [key j][64 channels] rows of 128
//     bytes (SWIZZLE_128B, 8-key groups of 1024 bytes, second 64-channel atom at LBO) -- no transposition pass;
//   * o accumulates in four 128-column TMEM chunks (C = 512 -> all 512 columns; the energies are dead by then);
//   * epilogue: TMEM -> y = act(gamma * o / rowsum + x) -> f32 and / or 16-bit planes.
// Scales (exact powers of two, undone in fp32): q, k x 16; v x 16; p x 1024 -- keep the lo halves out of fp16 subnormals.
// Shapes: Cq == 64, C % 128 == 0, C <= 512, HW <= 192 (the ShineOn U-Net: C = 512, HW = 12 / 48 / 192); anything else runs the
// CUDA-core kernels of attention.cu.
#include <cuda.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace shineon {

constexpr int kAtM = 128;       // query rows per CTA (UMMA M)
constexpr int kAtCq = 64;       // q / k channels (one 128-byte swizzle row)
constexpr int kAtCN = 128;      // channels per P.V accumulator chunk (UMMA N)
constexpr float kAtQKScale = 16.f, kAtVScale = 16.f, kAtPScale = 1024.f;

// MN-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::make_umma_desc<Major::MN>, LayoutType::B128):
// canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units = 128-byte rows along MN (64 elements), 8 K-rows per
// 1024-byte swizzle atom; LBO = byte distance between 64-element MN atoms, SBO = byte distance between 8-row K groups.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// 8 consecutive f32 -> one 16-byte chunk of the hi tile and of the lo tile (fp16, scaled)
__device__ __forceinline__ void at_store8(uint8_t* hi_chunk, uint8_t* lo_chunk, const float4 a, const float4 b, float scale) {
  uint32_t h[4], l[4];
  split16x2(a.x * scale, a.y * scale, SHINEON_FMT_FP16, h[0], l[0]);
  split16x2(a.z * scale, a.w * scale, SHINEON_FMT_FP16, h[1], l[1]);
  split16x2(b.x * scale, b.y * scale, SHINEON_FMT_FP16, h[2], l[2]);
  split16x2(b.z * scale, b.w * scale, SHINEON_FMT_FP16, h[3], l[3]);
  *reinterpret_cast<uint4*>(hi_chunk) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_chunk) = make_uint4(l[0], l[1], l[2], l[3]);
}

// smem map (from the 1024-aligned base):  region A = Q hi|lo, K hi|lo (phase 1)  ->  V chunk hi|lo (phase 3);  region P = P hi|lo
template <int HWP>  // keys padded to a multiple of 64 (64 or 192)
__global__ void __launch_bounds__(256, 1)
    sagan_attention_tc_kernel(const float* __restrict__ qkv, const float* __restrict__ x, const float* __restrict__ gamma,
                              float* __restrict__ yf, plane_t* __restrict__ yh, plane_t* __restrict__ yl, int HW, int C,
                              int cpad, int act, float act_param, int fmt) {
  pdl_grid_sync();
  constexpr int kKB = HWP / 64;                    // key k-blocks
  constexpr uint32_t kQBytes = kAtM * 128;         // one plane of Q
  constexpr uint32_t kKBytes = HWP * 128;          // one plane of K
  constexpr uint32_t kVAtom = HWP * 128;           // one 64-channel MN atom of a V chunk, all keys
  constexpr uint32_t kVPlane = 2 * kVAtom;         // 128 channels
  constexpr int kCPP = HWP == 64 ? 4 : 1;          // 128-channel V chunks staged (and multiplied) per pass: all of C for short key sets
  constexpr uint32_t kRegionA = (kCPP * 2 * kVPlane > 2 * (kQBytes + kKBytes)) ? kCPP * 2 * kVPlane : 2 * (kQBytes + kKBytes);
  constexpr uint32_t kPTile = kAtM * 128;          // one 64-key k-block of P, one plane
  constexpr uint32_t kPPlane = kKB * kPTile;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base_u = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (base_u - smem_u32(smem_raw));
  uint8_t* sA = base;                 // region A
  uint8_t* sP = base + kRegionA;      // P hi | P lo
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.y, i0 = blockIdx.x * kAtM;
  const int ld = 2 * kAtCq + C;
  const float* qkv_n = qkv + (long)n * HW * ld;
  const uint32_t bar_a = smem_u32(&bar);

  if (tid == 0) {
    mbar_init(bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(&tmem_slot));
  // ---- phase 1: Q (this tile's queries) and K (all keys) -> swizzled K-major fp16 hi/lo tiles
  {
    uint8_t* qh = sA;
    uint8_t* ql = sA + kQBytes;
    uint8_t* kh = sA + 2 * kQBytes;
    uint8_t* kl = kh + kKBytes;
    for (int it = tid; it < (kAtM + HWP) * 8; it += 256) {
      const int r = it >> 3, ch = it & 7;       // row, 16-byte chunk (8 channels)
      const bool isq = r < kAtM;
      const int row = isq ? r : r - kAtM;
      const int pix = isq ? i0 + row : row;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (pix < HW) {
        const float* src = qkv_n + (long)pix * ld + (isq ? 0 : kAtCq) + ch * 8;
        a = __ldg(reinterpret_cast<const float4*>(src));
        b = __ldg(reinterpret_cast<const float4*>(src + 4));
      }
      const uint32_t off = (uint32_t)row * 128u + (uint32_t)((ch ^ (row & 7)) * 16);
      at_store8((isq ? qh : kh) + off, (isq ? ql : kl) + off, a, b, kAtQKScale);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to tcgen05.mma
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t idesc_s = umma_idesc_f16(kAtM, HWP, 0);                    // A, B K-major, fp16, N = HWP
  const uint32_t idesc_o = umma_idesc_f16(kAtM, kAtCN, 0) | (1u << 16);     // B MN-major, N = 128
  uint32_t phase = 0;
  // ---- phase 2: energies S = Q K^T into TMEM columns [0, HWP)
  if (tid == 0) {
    const uint32_t qh = base_u, ql = base_u + kQBytes, kh = base_u + 2 * kQBytes, kl = kh + kKBytes;
#pragma unroll
    for (int k = 0; k < kAtCq / 16; ++k) {
      const uint64_t dQh = umma_desc_sw128(qh + k * 32), dQl = umma_desc_sw128(ql + k * 32);
      const uint64_t dKh = umma_desc_sw128(kh + k * 32), dKl = umma_desc_sw128(kl + k * 32);
      umma_f16(tmem_base, dQh, dKh, idesc_s, k != 0);
      umma_f16(tmem_base, dQh, dKl, idesc_s, 1);
      umma_f16(tmem_base, dQl, dKh, idesc_s, 1);
    }
    umma_commit(bar_a);
  }
  mbar_wait(bar_a, phase);
  phase ^= 1;
  tc_fence_after();
  // ---- softmax over the keys: thread = query row = TMEM lane (warps 0-3); P -> K-major swizzled A operand
  float rowsum = 1.f;
  if (warp < 4) {
    const int r = warp * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float inv_s = 1.f / (kAtQKScale * kAtQKScale);
    float m = -INFINITY;
#pragma unroll 1
    for (int c0 = 0; c0 < HWP; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(trow + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c0 + i < HW) m = fmaxf(m, __uint_as_float(v[i]));
    }
    float sum = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < HWP; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(trow + c0, v);
      tmem_ld_wait();
      float p[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        p[i] = (c0 + i < HW) ? expf((__uint_as_float(v[i]) - m) * inv_s) : 0.f;  // softmax numerator (sagan.py:45)
        sum += p[i];
      }
      uint8_t* ph = sP + (c0 >> 6) * kPTile + r * 128;
      uint8_t* pl = ph + kPPlane;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int chunk = (((c0 & 63) >> 3) + t) ^ (r & 7);
        at_store8(ph + chunk * 16, pl + chunk * 16, make_float4(p[8 * t], p[8 * t + 1], p[8 * t + 2], p[8 * t + 3]),
                  make_float4(p[8 * t + 4], p[8 * t + 5], p[8 * t + 6], p[8 * t + 7]), kAtPScale);
      }
    }
    rowsum = sum;
  }
  // ---- phase 3: o = P V, kCPP 128-channel chunks per pass (V staged as the MN-major B operand)
  const int n_chunks = C / kAtCN;
  for (int c_first = 0; c_first < n_chunks; c_first += kCPP) {
    if (c_first > 0) {  // the previous pass's MMAs have consumed region A
      mbar_wait(bar_a, phase);
      phase ^= 1;
    }
    const int n_pass = min(kCPP, n_chunks - c_first);
    for (int it = tid; it < HWP * 16 * n_pass; it += 256) {
      const int j = (it >> 4) % HWP, ch = it & 15, cl = it / (HWP * 16);  // key, 8-channel chunk of the 128-channel window, window
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (j < HW) {
        const float* src = qkv_n + (long)j * ld + 2 * kAtCq + (c_first + cl) * kAtCN + ch * 8;
        a = __ldg(reinterpret_cast<const float4*>(src));
        b = __ldg(reinterpret_cast<const float4*>(src + 4));
      }
      uint8_t* vh = sA + (uint32_t)cl * 2u * kVPlane;
      const uint32_t off = (uint32_t)(ch >> 3) * kVAtom + (uint32_t)j * 128u + (uint32_t)(((ch & 7) ^ (j & 7)) * 16);
      at_store8(vh + off, vh + kVPlane + off, a, b, kAtVScale);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();  // also orders the softmax warps' P stores (first pass) before the MMAs
    tc_fence_after();
    if (tid == 0) {
      const uint32_t ph = base_u + kRegionA, pl = ph + kPPlane;
#pragma unroll 1
      for (int cl = 0; cl < n_pass; ++cl) {
        const uint32_t vhh = base_u + (uint32_t)cl * 2u * kVPlane, vll = vhh + kVPlane;
        const uint32_t tacc = tmem_base + (uint32_t)((c_first + cl) * kAtCN);
#pragma unroll 1
        for (int kb = 0; kb < kKB; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t dPh = umma_desc_sw128(ph + kb * kPTile + k * 32), dPl = umma_desc_sw128(pl + kb * kPTile + k * 32);
            const uint32_t voff = (uint32_t)(kb * 8 + k * 2) * 1024u;  // 16 keys per MMA = two 8-key groups
            const uint64_t dVh = umma_desc_mn_sw128(vhh + voff, kVAtom, 1024u), dVl = umma_desc_mn_sw128(vll + voff, kVAtom, 1024u);
            umma_f16(tacc, dPh, dVh, idesc_o, (kb | k) != 0);
            umma_f16(tacc, dPh, dVl, idesc_o, 1);
            umma_f16(tacc, dPl, dVh, idesc_o, 1);
          }
        }
      }
      umma_commit(bar_a);
    }
  }
  mbar_wait(bar_a, phase);
  tc_fence_after();
  // ---- epilogue: y = act(gamma * o / rowsum + x); warp w: TMEM lanes of quarter w % 4, column half w / 4
  {
    const int q = warp & 3, half = warp >> 2;
    const int r = q * 32 + lane;
    const int pix_i = i0 + r;
    __shared__ float s_rowsum[kAtM];  // the softmax warps own the row sums; warps 4-7 read theirs from here
    if (warp < 4) s_rowsum[r] = rowsum;
    __syncthreads();
    const float g = __ldg(gamma);
    const float inv = g / (s_rowsum[r] * kAtPScale * kAtVScale);
    const bool ok = pix_i < HW;
    const long pix = (long)n * HW + pix_i;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool quarter_live = i0 + q * 32 < HW;  // warp-uniform: quarters past the last query have nothing to store
#pragma unroll 1
    for (int c0 = half * 32; quarter_live && c0 < C; c0 += 64) {
      uint32_t v[32];
      tmem_ld32(trow + c0, v);
      tmem_ld_wait();
      if (!ok) continue;
      float vals[32];
      const float4* xr = reinterpret_cast<const float4*>(x + pix * C + c0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 xv = __ldg(xr + i);
        vals[4 * i] = fmaf(__uint_as_float(v[4 * i]), inv, xv.x);  // sagan.py:53
        vals[4 * i + 1] = fmaf(__uint_as_float(v[4 * i + 1]), inv, xv.y);
        vals[4 * i + 2] = fmaf(__uint_as_float(v[4 * i + 2]), inv, xv.z);
        vals[4 * i + 3] = fmaf(__uint_as_float(v[4 * i + 3]), inv, xv.w);
      }
      switch (act) {  // warp-uniform
        case SHINEON_ACT_NONE: break;
        case SHINEON_ACT_GELU:
#pragma unroll
          for (int i = 0; i < 32; ++i) vals[i] = apply_act(vals[i], SHINEON_ACT_GELU, act_param);
          break;
        case SHINEON_ACT_RELU:
#pragma unroll
          for (int i = 0; i < 32; ++i) vals[i] = apply_act(vals[i], SHINEON_ACT_RELU, act_param);
          break;
        default:
#pragma unroll 1
          for (int i = 0; i < 32; ++i) vals[i] = apply_act(vals[i], act, act_param);
          break;
      }
      if (yf) {
        float4* dst = reinterpret_cast<float4*>(yf + pix * C + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = make_float4(vals[4 * i], vals[4 * i + 1], vals[4 * i + 2], vals[4 * i + 3]);
      }
      if (yh) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) split16x2(vals[8 * i + 2 * k], vals[8 * i + 2 * k + 1], fmt, h[k], l[k]);
          *reinterpret_cast<uint4*>(yh + pix * cpad + c0 + 8 * i) = make_uint4(h[0], h[1], h[2], h[3]);
          if (yl) *reinterpret_cast<uint4*>(yl + pix * cpad + c0 + 8 * i) = make_uint4(l[0], l[1], l[2], l[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace shineon

using namespace shineon;

// Returns SHINEON_OK after launching, a positive value when the shape does not fit this kernel (caller falls back).
int shineon_sagan_attention_tc(const float* qkv, const float* x, const float* gamma, float* y_f32, void* y_hi, void* y_lo,
                               int N, int HW, int C, int Cq, int cpad, int act, float act_param, int plane_fmt,
                               cudaStream_t stream) {
  if (Cq != kAtCq || C % kAtCN != 0 || C > 512 || HW > 192 || HW < 1) return 1;
  if (y_hi && cpad % 8 != 0) return 1;
  if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y_f32) |
       reinterpret_cast<uintptr_t>(y_hi) | reinterpret_cast<uintptr_t>(y_lo)) % 16 != 0)
    return 1;
  const int hwp = HW <= 64 ? 64 : 192;
  auto smem_of = [](int HWP) {
    const size_t qk = 2 * ((size_t)kAtM * 128 + (size_t)HWP * 128), v = (HWP == 64 ? 4 : 1) * 2 * (size_t)2 * HWP * 128;
    return (qk > v ? qk : v) + 2 * (size_t)(HWP / 64) * kAtM * 128 + 1024;
  };
  const size_t smem = smem_of(hwp);
  const dim3 grid(cdiv(HW, kAtM), N);
  cudaError_t e = cudaSuccess;
  if (hwp == 64) {
    static bool opted = false;
    if (!opted) {
      e = cudaFuncSetAttribute(sagan_attention_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      opted = e == cudaSuccess;
    }
    if (e == cudaSuccess)
      klaunch(sagan_attention_tc_kernel<64>, grid, 256, smem, stream, qkv, x, gamma, y_f32, (plane_t*)y_hi, (plane_t*)y_lo, HW, C, cpad,
                                                                 act, act_param, plane_fmt);
  } else {
    static bool opted = false;
    if (!opted) {
      e = cudaFuncSetAttribute(sagan_attention_tc_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      opted = e == cudaSuccess;
    }
    if (e == cudaSuccess)
      klaunch(sagan_attention_tc_kernel<192>, grid, 256, smem, stream, qkv, x, gamma, y_f32, (plane_t*)y_hi, (plane_t*)y_lo, HW, C, cpad,
                                                                  act, act_param, plane_fmt);
  }
  if (e != cudaSuccess) return fail(SHINEON_ERR_CUDA, "sagan_attention_tc: shared memory opt-in: %s", cudaGetErrorString(e));
  return after_launch("sagan_attention_tc_kernel");
}
