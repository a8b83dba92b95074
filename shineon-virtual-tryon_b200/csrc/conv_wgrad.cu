// Weight gradient of every Conv2d on the training path (rows U6 of SURVEY.md §8; the reference gets it from
// cuDNN through autograd: models/networks/cpvton/unet.py:129-174, attention/sagan.py:16-24) on tcgen05.
//
//   dW[co, (fy,fx), ci] = sum_{n,oh,ow} G[n,oh,ow,co] * X[n, oh*s+fy-p, ow*s+fx-p, ci]
//
// GEMM view per filter tap:  D[M = 128 co, N = BN ci] += G_tile^T[co, 64 px] * X_tile^T[ci, 64 px]^T, the
// contraction (K) runs over output pixels.  Both operands are NHWC 16-bit planes, i.e. CHANNEL-contiguous: the
// same TMA boxes the forward conv uses ({64 ch, bw, bh, nb}, shifted by the tap, zero-filled outside the image;
// the 5-D parity view for stride 2) land in shared memory as [pixel][64 ch] SWIZZLE_128B tiles, which is exactly
// the canonical *MN-major* UMMA operand layout (cute::UMMA::Major::MN, ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte
// units): no transposed copies of activations or gradients are ever written.
//
// Work item = (K-split, tap, co tile, ci tile); a persistent CTA walks items, accumulates in TMEM (double
// buffered so the epilogue overlaps the next item's main loop) and stores an f32 partial
// ws[split][co][tap][ci_pad]; wgrad_reduce_kernel sums the splits in a fixed order (deterministic) and scatters
// into the parameter's OIHW .grad.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = epilogue.
#include "common.cuh"
#include "tcgen05.cuh"

namespace shineon {

constexpr int kWgRows = 64;                     // pixels per K block (one TMA box = 64 rows x 128 B)
constexpr int kWgBoxBytes = kWgRows * 128;      // 8 KiB
constexpr int kWgABytes = 2 * kWgBoxBytes;      // M = 128 channels = two 64-channel boxes
constexpr int kWgThreads = 192;
constexpr int kWgMaxStages = 8;

struct WgradArgs {
  int nb, bh, bw, tiles_w, tiles_h;
  int stride, pad_h, pad_w, kw;
  int x_cstride;
  int Cout, g_cpad, g_coffset, cin_pad;
  int co_tiles, ci_tiles, taps, splits, kt_per_split, k_tiles;
  int total_items, stages;
  uint32_t idesc, lbo, sbo;
  float* ws;
};

// MN-major SWIZZLE_128B smem descriptor: LBO = byte distance between 64-element blocks along M/N,
// SBO = byte distance between 8-row groups along K.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN, bool SPLIT>
__global__ void __launch_bounds__(kWgThreads, 1)
    conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmGh, const __grid_constant__ CUtensorMap tmGl,
                      const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl,
                      const WgradArgs a) {
  pdl_grid_sync();
  constexpr int kBBoxes = BN / 64;
  constexpr int kBBytes = kBBoxes * kWgBoxBytes;
  constexpr int kPlanes = SPLIT ? 2 : 1;
  constexpr int kStageBytes = kPlanes * (kWgABytes + kBBytes);
  constexpr int kTmemCols = 2 * BN;  // double-buffered accumulators

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kWgMaxStages + 4];
  __shared__ uint32_t tmem_slot;

  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.stages;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[kWgMaxStages]),
                 bar_tfull = smem_u32(&bars[2 * kWgMaxStages]), bar_tempty = smem_u32(&bars[2 * kWgMaxStages + 2]);

  // Rows the TMA boxes never write (nb*bh*bw < 64) are K entries of every product: they must be zero, not stale.
  {
    uint4* z = reinterpret_cast<uint4*>(smem_raw + (tiles - smem_u32(smem_raw)));
    const int n16 = S * kStageBytes / 16;
    for (int i = threadIdx.x; i < n16; i += kWgThreads) z[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmGh);
    prefetch_tmap(&tmXh);
    if (SPLIT) {
      prefetch_tmap(&tmGl);
      prefetch_tmap(&tmXl);
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<kTmemCols>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int tiles_hw = a.tiles_w * a.tiles_h;
  const uint32_t box_bytes = (uint32_t)(a.nb * a.bh * a.bw) * 128u;

  // item -> (split, tap, co tile, ci tile); ci tile fastest: neighbouring CTAs share the G tile and the pixel range
  auto decode = [&](int item, int& split, int& tap, int& ct, int& it) {
    it = item % a.ci_tiles;
    int r = item / a.ci_tiles;
    ct = r % a.co_tiles;
    r /= a.co_tiles;
    tap = r % a.taps;
    split = r / a.taps;
  };

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      uint32_t g = 0;
      for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
        int split, tap, ct, it;
        decode(item, split, tap, ct, it);
        const int co0 = ct * 128, ci0 = it * BN;
        const int a_boxes = min(2, (a.g_cpad - co0) / 64);
        const int b_boxes = min(kBBoxes, (a.cin_pad - ci0) / 64);
        const uint32_t tx = kPlanes * (uint32_t)(a_boxes + b_boxes) * box_bytes;
        const int fy = tap / a.kw, fx = tap - fy * a.kw;
        const int kt0 = split * a.kt_per_split, kt1 = min(kt0 + a.kt_per_split, a.k_tiles);
        for (int kt = kt0; kt < kt1; ++kt, ++g) {
          const int s = g % S;
          const uint32_t ph = (g / S) & 1;
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          const uint32_t full = bar_full + 8 * s;
          mbar_expect_tx(full, tx);
          const int tw = kt % a.tiles_w, th = (kt / a.tiles_w) % a.tiles_h, tn = kt / tiles_hw;
          const int n0 = tn * a.nb, h0 = th * a.bh, w0 = tw * a.bw;
          const uint32_t sA = tiles + s * kStageBytes;
          const uint32_t sB = sA + kPlanes * kWgABytes;
          for (int b = 0; b < a_boxes; ++b) {
            tma_load_4d(sA + b * kWgBoxBytes, &tmGh, full, a.g_coffset + co0 + b * 64, w0, h0, n0);
            if (SPLIT) tma_load_4d(sA + kWgABytes + b * kWgBoxBytes, &tmGl, full, a.g_coffset + co0 + b * 64, w0, h0, n0);
          }
          if (a.stride == 1) {
            const int cx = w0 + fx - a.pad_w, cy = h0 + fy - a.pad_h;
            for (int b = 0; b < b_boxes; ++b) {
              tma_load_4d(sB + b * kWgBoxBytes, &tmXh, full, ci0 + b * 64, cx, cy, n0);
              if (SPLIT) tma_load_4d(sB + kBBytes + b * kWgBoxBytes, &tmXl, full, ci0 + b * 64, cx, cy, n0);
            }
          } else {
            const int ty = fy - a.pad_h, tx_ = fx - a.pad_w;
            const int py = ty & 1, px = tx_ & 1;
            const int ay = (ty - py) >> 1, ax = (tx_ - px) >> 1;
            for (int b = 0; b < b_boxes; ++b) {
              const int cc = px * a.x_cstride + ci0 + b * 64;
              tma_load_5d(sB + b * kWgBoxBytes, &tmXh, full, cc, w0 + ax, py, h0 + ay, n0);
              if (SPLIT) tma_load_5d(sB + kBBytes + b * kWgBoxBytes, &tmXl, full, cc, w0 + ax, py, h0 + ay, n0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (single thread)
    if (lane == 0) {
      uint32_t g = 0;
      int n_it = 0;
      for (int item = blockIdx.x; item < a.total_items; item += gridDim.x, ++n_it) {
        int split, tap, ct, it;
        decode(item, split, tap, ct, it);
        const int kt0 = split * a.kt_per_split, kt1 = min(kt0 + a.kt_per_split, a.k_tiles);
        const int buf = n_it & 1;
        mbar_wait(bar_tempty + 8 * buf, ((n_it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + buf * BN;
        for (int kt = kt0; kt < kt1; ++kt, ++g) {
          const int s = g % S;
          const uint32_t ph = (g / S) & 1;
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          const uint32_t sA = tiles + s * kStageBytes;
          const uint32_t sB = sA + kPlanes * kWgABytes;
#pragma unroll
          for (int k = 0; k < kWgRows / 16; ++k) {  // 16 pixels = two 8-row swizzle atoms = 2048 B
            const uint64_t dAh = umma_desc_mn_sw128(sA + k * 2048, a.lbo, a.sbo);
            const uint64_t dBh = umma_desc_mn_sw128(sB + k * 2048, a.lbo, a.sbo);
            umma_f16(tmem_acc, dAh, dBh, a.idesc, (kt != kt0) || (k != 0));
            if (SPLIT) {
              const uint64_t dAl = umma_desc_mn_sw128(sA + kWgABytes + k * 2048, a.lbo, a.sbo);
              const uint64_t dBl = umma_desc_mn_sw128(sB + kBBytes + k * 2048, a.lbo, a.sbo);
              umma_f16(tmem_acc, dAh, dBl, a.idesc, 1);
              umma_f16(tmem_acc, dAl, dBh, a.idesc, 1);
            }
          }
          umma_commit(bar_empty + 8 * s);
        }
        umma_commit(bar_tfull + 8 * buf);
      }
    }
  } else {
    // ===================================================== epilogue: TMEM -> registers -> f32 partials
    const int q = warp & 3;
    int n_it = 0;
    for (int item = blockIdx.x; item < a.total_items; item += gridDim.x, ++n_it) {
      int split, tap, ct, it;
      decode(item, split, tap, ct, it);
      const int co = ct * 128 + q * 32 + lane;
      const int ci0 = it * BN;
      const int buf = n_it & 1;
      mbar_wait(bar_tfull + 8 * buf, (n_it >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + buf * BN;
      float* dst = a.ws + (((long)split * a.Cout + co) * a.taps + tap) * a.cin_pad + ci0;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (ci0 + c0 >= a.cin_pad) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        if (co < a.Cout) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]),
                                                                   __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

// Sums the K-split partials (fixed order) and scatters into the parameter gradient.
//   mode 0: grad is OIHW [Cout][Cin][kh][kw]; packed channel cp -> ci through chan_map (or identity for cp < Cin)
//   mode 1: im2col'd first layer: the GEMM ran as a 1x1 conv over k = tap*Cin + ci (taps == 1, cin_pad == kpad)
//   mode 2: grad is ConvTranspose2d IOHW [Cin][Cout][kh][kw] with flipped taps (the packed weight of a deconv)
__global__ void __launch_bounds__(256)
    wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ grad, const int32_t* __restrict__ chan_map,
                        int splits, int Cout, int taps, int cin_pad, int Cin, int kh, int kw, int mode, float alpha,
                        float beta) {
  pdl_grid_sync();
  const long split_stride = (long)Cout * taps * cin_pad;
  if (mode == 1) {  // im2col'd first layer: small, one element per thread
    const long total = (long)Cout * cin_pad;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
      const int cp = (int)(e % cin_pad);
      const int co = (int)(e / cin_pad);
      if (cp >= kh * kw * Cin) continue;
      const int t = cp / Cin, ci = cp - t * Cin;
      const long o = ((long)co * Cin + ci) * kh * kw + t;
      float s = 0.f;
      for (int k = 0; k < splits; ++k) s += ws[k * split_stride + e];
      grad[o] = beta == 0.f ? alpha * s : fmaf(beta, grad[o], alpha * s);
    }
    return;
  }
  // one thread per (co, packed channel): reads are coalesced over the channel for every tap, the taps of one
  // (co, ci) are contiguous in the OIHW gradient
  const long total = (long)Cout * cin_pad;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int cp = (int)(e % cin_pad);
    const int co = (int)(e / cin_pad);
    const int ci = chan_map ? chan_map[cp] : (cp < Cin ? cp : -1);
    if (ci < 0 || ci >= Cin) continue;
    const float* src = ws + ((long)co * taps) * cin_pad + cp;
    for (int tap = 0; tap < taps; ++tap) {
      float s = 0.f;
      for (int k = 0; k < splits; ++k) s += src[k * split_stride + (long)tap * cin_pad];
      long o;
      if (mode == 0) {
        o = ((long)co * Cin + ci) * taps + tap;
      } else {
        const int fy = tap / kw, fx = tap - fy * kw;
        o = (((long)ci * Cout + co) * kh + (kh - 1 - fy)) * kw + (kw - 1 - fx);
      }
      grad[o] = beta == 0.f ? alpha * s : fmaf(beta, grad[o], alpha * s);
    }
  }
}

// Per-channel sum over all pixels of an NHWC f32 tensor (bias gradients): grad[c] = beta*grad[c] + alpha*sum.
__global__ void __launch_bounds__(256)
    channel_sum_kernel(const float* __restrict__ x, double* __restrict__ acc, long pixels, int C, int cstride,
                       int pix_per_cta) {
  pdl_grid_sync();
  // block: 32 channels x 8 pixel lanes
  __shared__ float s[8][33];
  const int c = blockIdx.y * 32 + (threadIdx.x & 31);
  const int pl = threadIdx.x >> 5;
  const long p0 = (long)blockIdx.x * pix_per_cta, p1 = min(p0 + (long)pix_per_cta, pixels);
  float v = 0.f;
  if (c < C)
    for (long p = p0 + pl; p < p1; p += 8) v += __ldg(x + p * cstride + c);
  s[pl][threadIdx.x & 31] = v;
  __syncthreads();
  if (pl == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s[i][threadIdx.x];
    atomicAdd(acc + c, (double)t);
  }
}
__global__ void channel_sum_finish_kernel(const double* __restrict__ acc, float* __restrict__ grad, int C, float alpha,
                                          float beta) {
  pdl_grid_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) grad[c] = beta == 0.f ? alpha * (float)acc[c] : fmaf(beta, grad[c], alpha * (float)acc[c]);
}

static void pick_tile_rows(int N, int Ho, int Wo, int rows, int& nb, int& bh, int& bw) {
  double best = -1.0;
  nb = 1; bh = 1; bw = 1;
  for (int w = 1; w <= Wo && w <= rows; ++w) {
    for (int h = 1; h <= Ho && h * w <= rows; ++h) {
      int n = rows / (w * h);
      if (n > N) n = N;
      if (n < 1) n = 1;
      long tiles = (long)cdiv(Wo, w) * cdiv(Ho, h) * cdiv(N, n);
      double eff = (double)N * Ho * Wo / ((double)tiles * rows) + 1e-6 * w;
      if (eff > best) { best = eff; nb = n; bh = h; bw = w; }
    }
  }
}

template <int BN, bool SPLIT>
static int launch_wgrad(const CUtensorMap& tGh, const CUtensorMap& tGl, const CUtensorMap& tXh, const CUtensorMap& tXl,
                        WgradArgs& a, int fmt, cudaStream_t stream) {
  constexpr int kStageBytes = (SPLIT ? 2 : 1) * (kWgABytes + (BN / 64) * kWgBoxBytes);
  static int max_dyn_smem = -1;
  if (max_dyn_smem < 0) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, conv_wgrad_kernel<BN, SPLIT>);
    if (e == cudaSuccess) {
      max_dyn_smem = 227 * 1024 - (int)fa.sharedSizeBytes;
      e = cudaFuncSetAttribute(conv_wgrad_kernel<BN, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn_smem);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      max_dyn_smem = -1;
      return fail(SHINEON_ERR_CUDA, "conv_wgrad shared-memory opt-in: %s", cudaGetErrorString(e));
    }
  }
  int stages = (max_dyn_smem - 1024) / kStageBytes;
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  if (stages > a.kt_per_split + 1) stages = a.kt_per_split + 1;
  if (stages < 1) return fail(SHINEON_ERR_UNSUPPORTED, "conv_wgrad: tile does not fit in shared memory");
  a.stages = stages;
  // instruction descriptor: D = f32, A/B = fp16|bf16, both MN-major (bits 15, 16), N = BN, M = 128
  a.idesc = umma_idesc_f16(128, BN, fmt == SHINEON_FMT_FP16 ? 0 : 1) | (1u << 15) | (1u << 16);
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
  }
  const int grid = a.total_items < num_sms ? a.total_items : num_sms;
  klaunch(conv_wgrad_kernel<BN, SPLIT>, grid, kWgThreads, stages * kStageBytes + 1024, stream, tGh, tGl, tXh, tXl, a);
  return after_launch("conv_wgrad_kernel");
}

}  // namespace shineon

using namespace shineon;

extern "C" size_t shineon_conv2d_wgrad_workspace_bytes(const shineon_conv2d_wgrad_params* p) {
  if (!p || p->N <= 0 || p->Ho <= 0 || p->Wo <= 0 || p->Cout <= 0 || p->cin_pad <= 0) return 0;
  int nb, bh, bw;
  pick_tile_rows(p->N, p->Ho, p->Wo, kWgRows, nb, bh, bw);
  const int k_tiles = cdiv(p->Wo, bw) * cdiv(p->Ho, bh) * cdiv(p->N, nb);
  const int bn = p->cin_pad >= 128 ? 128 : 64;
  const int taps = p->mode == 1 ? 1 : p->kh * p->kw;
  const long base = (long)taps * cdiv(p->Cout, 128) * cdiv(p->cin_pad, bn);
  int splits = (int)((2 * 148 + base - 1) / base);
  if (p->splits > 0) splits = p->splits;
  if (splits > k_tiles) splits = k_tiles;
  if (splits < 1) splits = 1;
  const int per = cdiv(k_tiles, splits);
  splits = cdiv(k_tiles, per);
  return (size_t)splits * p->Cout * taps * p->cin_pad * sizeof(float) + 256;
}

extern "C" int shineon_conv2d_wgrad(const shineon_conv2d_wgrad_params* p, shineon_stream_t stream_) {
  SHINEON_REQUIRE(p != nullptr, "conv2d_wgrad: null params");
  SHINEON_REQUIRE(p->g_hi && p->x_hi && p->grad_w && p->workspace, "conv2d_wgrad: null pointer");
  SHINEON_REQUIRE((p->g_lo == nullptr) == (p->x_lo == nullptr), "conv2d_wgrad: g_lo and x_lo must both be given or both NULL");
  SHINEON_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0 && p->Ho > 0 && p->Wo > 0 && p->Cout > 0 && p->Cin > 0, "conv2d_wgrad: bad shape");
  SHINEON_REQUIRE(p->cin_pad % 64 == 0 && p->g_cpad % 64 == 0 && p->g_cpad >= 64, "conv2d_wgrad: channel padding must be a multiple of 64");
  SHINEON_REQUIRE(p->kh > 0 && p->kw > 0 && p->kh <= 7 && p->kw <= 7, "conv2d_wgrad: kernel %dx%d unsupported", p->kh, p->kw);
  SHINEON_REQUIRE(p->stride == 1 || (p->stride == 2 && p->H % 2 == 0 && p->W % 2 == 0), "conv2d_wgrad: stride %d", p->stride);
  SHINEON_REQUIRE(p->plane_fmt == SHINEON_FMT_BF16 || p->plane_fmt == SHINEON_FMT_FP16, "conv2d_wgrad: plane_fmt %d", p->plane_fmt);
  SHINEON_REQUIRE(p->mode >= 0 && p->mode <= 2, "conv2d_wgrad: mode %d", p->mode);
  const int x_cstride = p->x_cstride ? p->x_cstride : p->cin_pad;
  const int g_cstride = p->g_cstride ? p->g_cstride : p->g_cpad;
  SHINEON_REQUIRE(p->g_coffset >= 0 && p->g_coffset + p->Cout <= g_cstride, "conv2d_wgrad: G channel window out of range");
  SHINEON_REQUIRE(x_cstride >= p->cin_pad && x_cstride % 8 == 0 && g_cstride >= p->g_cpad && g_cstride % 8 == 0, "conv2d_wgrad: channel strides");
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool split = p->g_lo != nullptr;

  WgradArgs a;
  pick_tile_rows(p->N, p->Ho, p->Wo, kWgRows, a.nb, a.bh, a.bw);
  a.tiles_w = cdiv(p->Wo, a.bw);
  a.tiles_h = cdiv(p->Ho, a.bh);
  a.k_tiles = a.tiles_w * a.tiles_h * cdiv(p->N, a.nb);
  // mode 1: the GEMM is a 1x1 conv over the im2col'd K; kh/kw/Cin only describe the parameter for the scatter
  const bool i2c = p->mode == 1;
  a.stride = i2c ? 1 : p->stride; a.pad_h = i2c ? 0 : p->pad_h; a.pad_w = i2c ? 0 : p->pad_w; a.kw = i2c ? 1 : p->kw;
  a.x_cstride = x_cstride;
  a.Cout = p->Cout; a.g_cpad = p->g_cpad; a.g_coffset = p->g_coffset; a.cin_pad = p->cin_pad;
  const int bn = p->cin_pad >= 128 ? 128 : 64;
  a.co_tiles = cdiv(p->Cout, 128);
  a.ci_tiles = cdiv(p->cin_pad, bn);
  a.taps = i2c ? 1 : p->kh * p->kw;
  const long base = (long)a.taps * a.co_tiles * a.ci_tiles;
  int splits = (int)((2 * 148 + base - 1) / base);
  if (p->splits > 0) splits = p->splits;
  if (splits > a.k_tiles) splits = a.k_tiles;
  if (splits < 1) splits = 1;
  a.kt_per_split = cdiv(a.k_tiles, splits);
  a.splits = cdiv(a.k_tiles, a.kt_per_split);  // no empty split
  const long total = base * a.splits;
  SHINEON_REQUIRE(total < (1l << 31), "conv2d_wgrad: too many work items");
  a.total_items = (int)total;
  const size_t need = (size_t)a.splits * p->Cout * a.taps * p->cin_pad * sizeof(float);
  SHINEON_REQUIRE(p->workspace_bytes >= need, "conv2d_wgrad: workspace %zu < %zu bytes", p->workspace_bytes, need);
  a.ws = (float*)p->workspace;
  // descriptor strides; desc_variant 1 swaps the two fields (kept for bring-up diagnostics)
  a.lbo = kWgBoxBytes;
  a.sbo = 1024;
  if (p->desc_variant == 1) { a.lbo = 1024; a.sbo = kWgBoxBytes; }

  int rc;
  CUtensorMap tGh, tGl, tXh, tXl;
  {
    const cuuint64_t C = (cuuint64_t)g_cstride;
    // the channel extent covers the whole pixel row: a window (g_coffset) may start anywhere inside it
    cuuint64_t dims[4] = {C, (cuuint64_t)p->Wo, (cuuint64_t)p->Ho, (cuuint64_t)p->N};
    cuuint64_t strides[3] = {C * 2, (cuuint64_t)p->Wo * C * 2, (cuuint64_t)p->Ho * p->Wo * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)a.bw, (cuuint32_t)a.bh, (cuuint32_t)a.nb};
    if ((rc = encode_map(&tGh, p->g_hi, 4, dims, strides, box, "G hi", p->plane_fmt))) return rc;
    if (split && (rc = encode_map(&tGl, p->g_lo, 4, dims, strides, box, "G lo", p->plane_fmt))) return rc;
  }
  {
    const cuuint64_t Cw = (cuuint64_t)p->cin_pad, C = (cuuint64_t)x_cstride, H = (cuuint64_t)p->H, W = (cuuint64_t)p->W, N = (cuuint64_t)p->N;
    if (a.stride == 1) {
      cuuint64_t dims[4] = {Cw, W, H, N};
      cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
      cuuint32_t box[4] = {64, (cuuint32_t)a.bw, (cuuint32_t)a.bh, (cuuint32_t)a.nb};
      if ((rc = encode_map(&tXh, p->x_hi, 4, dims, strides, box, "X hi", p->plane_fmt))) return rc;
      if (split && (rc = encode_map(&tXl, p->x_lo, 4, dims, strides, box, "X lo", p->plane_fmt))) return rc;
    } else {
      cuuint64_t dims[5] = {C + Cw, W / 2, 2, H / 2, N};
      cuuint64_t strides[4] = {2 * C * 2, W * C * 2, 2 * W * C * 2, H * W * C * 2};
      cuuint32_t box[5] = {64, (cuuint32_t)a.bw, 1, (cuuint32_t)a.bh, (cuuint32_t)a.nb};
      if ((rc = encode_map(&tXh, p->x_hi, 5, dims, strides, box, "X hi s2", p->plane_fmt))) return rc;
      if (split && (rc = encode_map(&tXl, p->x_lo, 5, dims, strides, box, "X lo s2", p->plane_fmt))) return rc;
    }
  }
  if (!split) { tGl = tGh; tXl = tXh; }
  if (bn == 128)
    rc = split ? launch_wgrad<128, true>(tGh, tGl, tXh, tXl, a, p->plane_fmt, stream)
               : launch_wgrad<128, false>(tGh, tGl, tXh, tXl, a, p->plane_fmt, stream);
  else
    rc = split ? launch_wgrad<64, true>(tGh, tGl, tXh, tXl, a, p->plane_fmt, stream)
               : launch_wgrad<64, false>(tGh, tGl, tXh, tXl, a, p->plane_fmt, stream);
  if (rc) return rc;
  {
    const long total_e = (long)p->Cout * p->cin_pad;
    long blocks = (total_e + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    klaunch(wgrad_reduce_kernel, (int)blocks, 256, 0, stream, a.ws, p->grad_w, p->chan_map, a.splits, p->Cout, a.taps, p->cin_pad,
                                                         p->Cin, p->kh, p->kw, p->mode, p->alpha == 0.f ? 1.f : p->alpha, p->beta);
    return after_launch("wgrad_reduce_kernel");
  }
}

extern "C" int shineon_channel_sum(const float* x, float* grad, void* workspace, long pixels, int C, int cstride,
                                   float alpha, float beta, shineon_stream_t stream_) {
  SHINEON_REQUIRE(x && grad && workspace && pixels > 0 && C > 0 && cstride >= C, "channel_sum: bad arguments");
  cudaStream_t stream = (cudaStream_t)stream_;
  cudaError_t e = cudaMemsetAsync(workspace, 0, sizeof(double) * C, stream);
  if (e != cudaSuccess) return fail(SHINEON_ERR_CUDA, "channel_sum memset: %s", cudaGetErrorString(e));
  int ctas = (int)((pixels + 255) / 256);
  if (ctas > 148 * 4) ctas = 148 * 4;
  const int per = (int)((pixels + ctas - 1) / ctas);
  ctas = (int)((pixels + per - 1) / per);
  klaunch(channel_sum_kernel, dim3(ctas, cdiv(C, 32)), 256, 0, stream, x, (double*)workspace, pixels, C, cstride, per);
  int rc = after_launch("channel_sum_kernel");
  if (rc) return rc;
  klaunch(channel_sum_finish_kernel, cdiv(C, 128), 128, 0, stream, (const double*)workspace, grad, C, alpha == 0.f ? 1.f : alpha, beta);
  return after_launch("channel_sum_finish_kernel");
}
