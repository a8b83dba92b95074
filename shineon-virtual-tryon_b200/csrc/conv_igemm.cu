// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Replaces every nn.Conv2d / nn.ConvTranspose2d on the hot path (cuDNN calls in the reference):
//   models/networks/cpvton/warp.py:14-28,74-84   (GMM 4x4 s2 / 3x3 s1)
//   models/networks/cpvton/unet.py:129-174       (U-Net 4x4 s2 down, 3x3 s1 up)
//   models/networks/attention/sagan.py:16-24     (1x1 q/k/v projections)
//   models/flownet2_pytorch/networks/submodules.py:7-38 (FlowNet2 7x7/5x5/3x3/1x1, deconv 4x4 s2)
//
// GEMM view:  D[M = 128 output pixels, N = BN output channels] += A_tap[M, 64 ch] * W_tap[N, 64 ch]^T
// over K-blocks kb = (tap, 64-channel slice).
//   * A tiles come straight from the NHWC activation with one TMA box per K-block: a 4-D box
//     {64 ch, bw, bh, nb} shifted by the filter tap for stride 1, and for stride 2 a 5-D box over the
//     [N, H/2, 2, W/2, 2*C] view of the same tensor (row/column parity folded into the coordinates).
//     Out-of-bounds box elements are zero-filled by TMA = the convolution's zero padding.
//   * B tiles are {64, BN} boxes of the packed weight matrix [Cout][taps*Cin_pad].
//   * Both land in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma consumes.
//   * One elected thread issues tcgen05.mma (M=128, N=BN, K=16, bf16 -> fp32 in TMEM); with hi/lo
//     planes it issues hi*hi + hi*lo + lo*hi (bf16x3: fp32-grade products).
//   * 4 epilogue warps read TMEM (tcgen05.ld 32x32b), apply bias / activation / per-channel affine
//     (folded BatchNorm) and store fp32 NHWC and/or bf16 hi/lo planes for the next layer.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = epilogue.
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "tcgen05.cuh"

namespace shineon {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                 // bf16 elements = one 128-byte swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kEpiWarps = 8;  // two per TMEM lane quarter (alternating 32-column chunks): one warp per scheduler was
                              // latency-bound on small-K layers (tcgen05.ld -> math -> store chain fully exposed)
constexpr int kThreads = (2 + kEpiWarps) * 32;

struct ConvArgs {
  // tile geometry
  int nb, bh, bw;
  int tiles_w, tiles_h;
  int N, Ho, Wo, Cout;
  int kh, kw, stride, pad_h, pad_w;
  int cin_pad, cin_blocks, x_cstride;
  int num_kb, stages;
  int chunk_kb;      // K-blocks accumulated in one TMEM buffer before the epilogue warps take the partial sum over
  int stage_off;     // byte offset (from the 1024-aligned tile base) of the epilogue transpose buffers, < 0 = direct stores
  int n_tiles, total_tiles;  // channel tiles per pixel tile, pixel tiles * channel tiles
  int w_per_image;           // 1: the B operand is a [N][Cout][K] batch indexed by the tile's image (needs nb == 1)
  // ConvTranspose2d(4, 2, 1) as ONE launch: the four output-parity phases (py, px) are an extra tile dimension (tile order:
  // channel tile, then phase, then pixel tile -- with the phase slowest, FlowNetFusion's deconv0 re-read its 201 MB input from
  // DRAM once per phase: 600 MB per launch under ncu, profiles/r02_late_kernels.md).
  // Phase tiles read the weight set [phase] of a [4][Cout][2*2][cin_pad] batch, pad (1 - py, 1 - px) and write the output
  // lattice (2*oh + py, 2*ow + px).  (Four launches of 8-32 CTAs each left the small pyramid levels of FlowNet2 and the
  // training step's down-conv input gradients on a sliver of the chip.)
  int phases;                // 1 or 4
  int phase_tiles;           // pixel tiles of one phase
  // epilogue
  const float* bias;
  const float* scale;
  const float* shift;
  int pre_act, post_act;
  float act_param;
  float acc_scale;   // multiplies the accumulator first (undoes the power-of-two weight scaling of fp16 planes)
  int fmt;           // plane encoding (SHINEON_FMT_*)
  uint32_t idesc;    // tcgen05 instruction descriptor for (fmt, BN)
  uint32_t idesc_cat;  // same with N = 2*BN (split mode: [B_hi | B_lo] in one MMA)
  float* y_f32;
  plane_t* y_hi;
  plane_t* y_lo;
  int out_H, out_W, out_cstride, out_coffset;
  int oh_mul, oh_off, ow_mul, ow_off;
  double* stats;  // [N][Cout][2] per-(image, channel) sum / sum of squares of the f32 output (InstanceNorm), nb == 1 only
  // split-K (layers with fewer tiles than SMs and a deep K: the 4x3 ... 8x6 levels): work item = (tile, K slice); every
  // item writes its raw partial accumulator to its own workspace slice, the LAST item of a tile to arrive (per-tile counter)
  // sums the slices in slice order -- deterministic -- and runs the normal epilogue
  int ksplit, kb_per_split;  // ksplit == 1: off
  float* sk_ws;              // [ksplit][m_tiles * 128][sk_ctot]
  int* sk_cnt;               // [total_tiles], zero before the launch; the finalising item resets its counter
  int sk_ctot;               // n_tiles * BN
  long sk_slice;             // floats per slice
};

// ------------------------------------------------------------------------------------- epilogue
// Scalar form (CUDA-core cross-check kernel).
__device__ __forceinline__ float conv_epilogue_value(float acc, int c, const ConvArgs& a) {
  float v = acc * a.acc_scale;
  if (a.bias) v += __ldg(a.bias + c);
  v = apply_act(v, a.pre_act, a.act_param);
  if (a.scale) v = fmaf(v, __ldg(a.scale + c), __ldg(a.shift + c));
  v = apply_act(v, a.post_act, a.act_param);
  return v;
}

// Vector form for the tcgen05 epilogue: the activation kind is warp-uniform, so dispatch ONCE per 32-column chunk
// to a loop specialised on it.  (A per-element `switch` unrolled 32x made the kernel 16k SASS instructions and
// instruction-fetch bound: stall_no_inst dominated the first ncu capture, profiles/r01_conv_igemm.md capture A.)
template <int ACT>
__device__ __forceinline__ void act_chunk(float (&v)[32], float param) {
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = apply_act(v[i], ACT, param);
}
__device__ __forceinline__ void act_chunk_dispatch(float (&v)[32], int act, float param) {
  switch (act) {  // warp-uniform
    case SHINEON_ACT_NONE: break;
    case SHINEON_ACT_RELU: act_chunk<SHINEON_ACT_RELU>(v, param); break;
    case SHINEON_ACT_LEAKY: act_chunk<SHINEON_ACT_LEAKY>(v, param); break;
    case SHINEON_ACT_GELU: act_chunk<SHINEON_ACT_GELU>(v, param); break;
    case SHINEON_ACT_SWISH: act_chunk<SHINEON_ACT_SWISH>(v, param); break;
    case SHINEON_ACT_SINE: act_chunk<SHINEON_ACT_SINE>(v, param); break;
    case SHINEON_ACT_TANH: act_chunk<SHINEON_ACT_TANH>(v, param); break;
    default: act_chunk<SHINEON_ACT_SIGMOID>(v, param); break;
  }
}

// Stores `cnt` (<= 32) consecutive channels of one pixel; the vector path needs cnt == 32 (or a multiple of 8
// for planes / 4 for f32) and an aligned base, which every 64-padded layer satisfies.
template <int FMT>
__device__ __forceinline__ void store_planes_vec(const float (&vals)[32], int cnt, plane_t* yh, plane_t* yl, long base) {
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    if (i < cnt) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split16x2(vals[i + 2 * j], vals[i + 2 * j + 1], FMT, hi[j], lo[j]);
      *reinterpret_cast<uint4*>(yh + base + i) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      if (yl) *reinterpret_cast<uint4*>(yl + base + i) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

__device__ __forceinline__ void conv_store_row(const float (&vals)[32], int c_first, int cnt, long pix, const ConvArgs& a) {
  const long base = pix * a.out_cstride + a.out_coffset + c_first;
  if (a.y_f32) {
    float* dst = a.y_f32 + base;
    if ((cnt & 3) == 0 && (base & 3) == 0) {
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        if (i < cnt) *reinterpret_cast<float4*>(dst + i) = make_float4(vals[i], vals[i + 1], vals[i + 2], vals[i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < cnt) dst[i] = vals[i];
    }
  }
  if (a.y_hi) {
    if ((cnt & 7) == 0 && (base & 7) == 0) {
      if (a.fmt == SHINEON_FMT_FP16)
        store_planes_vec<SHINEON_FMT_FP16>(vals, cnt, a.y_hi, a.y_lo, base);
      else
        store_planes_vec<SHINEON_FMT_BF16>(vals, cnt, a.y_hi, a.y_lo, base);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < cnt) {
          plane_t h, l;
          split16(vals[i], a.fmt, h, l);
          a.y_hi[base + i] = h;
          if (a.y_lo) a.y_lo[base + i] = l;
        }
    }
  }
}

// Coalesced form of conv_store_row for a full-warp 32-row x 32-channel chunk.  tcgen05.ld leaves each lane with ONE
// pixel row (32 consecutive channels = 128 B), so a direct 16-byte store instruction touches 32 different cache lines:
// the epilogue warps then sit on the store queue (ncu: half of all stall samples were the WAR wait on the address
// registers of in-flight stores, profiles/r01_conv_epilogue_store.md) and small-K layers are epilogue-bound.  The chunk
// is transposed through a 4 KB per-warp staging buffer (16-byte unit j of row r at j ^ (r & 7): conflict-free both
// ways) so that eight lanes write one pixel's 128 contiguous bytes and a store instruction covers 4 full lines.
// mybase = element offset (pix * out_cstride + out_coffset) of this lane's row; okmask bit r = row r is stored.
__device__ __forceinline__ void conv_store_chunk_coalesced(const float (&vals)[32], float4* stage, int lane,
                                                           long mybase, uint32_t okmask, int c_first, int cnt,
                                                           const ConvArgs& a) {
  // every lane needs the output offsets of the 8 rows it will write (row = it*4 + lane/8); fetched here, per chunk,
  // instead of living in 16 registers across the whole tile
  long rowbase[8];
#pragma unroll
  for (int it2 = 0; it2 < 8; ++it2) rowbase[it2] = __shfl_sync(0xffffffffu, mybase, it2 * 4 + (lane >> 3));
  __syncwarp();  // the previous chunk's reads of the staging buffer are done
#pragma unroll
  for (int j = 0; j < 8; ++j)
    stage[lane * 8 + (j ^ (lane & 7))] = make_float4(vals[4 * j], vals[4 * j + 1], vals[4 * j + 2], vals[4 * j + 3]);
  __syncwarp();
  const int u = lane & 7;
  if (4 * u >= cnt) return;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int row = it * 4 + (lane >> 3);
    if (!((okmask >> row) & 1u)) continue;
    const float4 v = stage[row * 8 + (u ^ (row & 7))];
    const long base = rowbase[it] + c_first + 4 * u;
    if (a.y_f32) *reinterpret_cast<float4*>(a.y_f32 + base) = v;
    if (a.y_hi) {
      uint32_t h01, h23, l01, l23;
      split16x2(v.x, v.y, a.fmt, h01, l01);
      split16x2(v.z, v.w, a.fmt, h23, l23);
      *reinterpret_cast<uint2*>(a.y_hi + base) = make_uint2(h01, h23);
      if (a.y_lo) *reinterpret_cast<uint2*>(a.y_lo + base) = make_uint2(l01, l23);
    }
  }
}

// -------------------------------------------------------------------------------------- kernel
template <int NTHREADS>  // named barrier 1 over the kernel's epilogue warps only
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory"); }

// Persistent: one CTA per SM walks output tiles (tile = blockIdx.x + i*gridDim.x, channel tile fastest so that
// neighbouring CTAs share the same activation tile in L2).  Two TMEM accumulators: the epilogue of tile i
// (TMEM -> registers -> global) overlaps the TMA/MMA main loop of tile i+1.
// MULTI: the K loop spans several accumulation chunks (num_kb > chunk_kb); SPLITK: work items are (tile, K slice) pairs
// (compiled separately so the common single-pass epilogue keeps its register budget)
template <int BN, bool SPLIT, bool MULTI, bool SPLITK = false>
__global__ void __launch_bounds__(kThreads, 1)
    conv_igemm_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                      const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                      const ConvArgs a) {
  pdl_grid_sync();
  constexpr int kBBytes = BN * kBlockK * 2;
  constexpr int kPlanes = SPLIT ? 2 : 1;
  constexpr int kStageBytes = kPlanes * (kABytes + kBBytes);
  // Split (hi/lo) mode issues hi*hi + hi*lo + lo*hi.  The B_hi and B_lo tiles sit back to back in shared memory, so
  // A_hi * [B_hi | B_lo] is ONE MMA with N = 2*BN into columns [0, 2*BN) (left half hi*hi, right half hi*lo) and
  // A_lo * B_hi a second one into the left half; the epilogue adds the halves.  Two MMAs instead of three and A_hi is
  // read from shared memory once instead of twice: the narrow-N (Cout <= 64) layers are bound by exactly that
  // operand traffic (128 B/clk/SM), profiles/r01_conv_igemm.md.  Needs 4*BN TMEM columns, i.e. BN <= 128.
  constexpr bool kCat = SPLIT && BN <= 128;
  constexpr int kAccCols = (kCat ? 2 : 1) * BN < 32 ? 32 : (kCat ? 2 : 1) * BN;  // TMEM columns per accumulator
  constexpr int kTmemCols = 2 * kAccCols;      // double-buffered (<= 512)
  const uint32_t kIdesc = a.idesc;
  const uint32_t kIdescCat = a.idesc_cat;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 4];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[BN], s_scale[BN], s_shift[BN];  // current tile's output-channel window
  // current tile's per-channel sum / sum of squares (a.stats), one slot per TMEM lane quarter: every slot has exactly one
  // writer warp and the flush adds the four in a fixed order, so the statistics are reproducible run to run
  __shared__ float s_stat[4][BN][2];
  __shared__ int s_flag;  // split-K: arrival order of this item among its tile's K slices

  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-B alignment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.stages;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[kMaxStages]),
                 bar_tfull = smem_u32(&bars[2 * kMaxStages]), bar_tempty = smem_u32(&bars[2 * kMaxStages + 2]);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmAh);
    prefetch_tmap(&tmBh);
    if (SPLIT) {
      prefetch_tmap(&tmAl);
      prefetch_tmap(&tmBl);
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);   // one tcgen05.commit
      mbar_init(bar_tempty + 8 * b, kEpiWarps);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<kTmemCols>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int tiles_hw = a.tiles_w * a.tiles_h;
  const int total_items = a.total_tiles * a.ksplit;  // work items: (tile, K slice), the slices of a tile on neighbouring CTAs

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      const uint32_t a_rows = a.nb * a.bh * a.bw;
      const uint32_t tx = kPlanes * (a_rows * (kBlockK * 2) + kBBytes);
      uint32_t g = 0;  // running K-block counter across tiles (ring position)
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int tile = item / a.ksplit, ks = item - tile * a.ksplit;
        const int kb_lo = ks * a.kb_per_split, kb_hi = min(a.num_kb, kb_lo + a.kb_per_split);
        const int nt = tile % a.n_tiles;
        int mt = tile / a.n_tiles;
        int pad_h = a.pad_h, pad_w = a.pad_w, phz = 0;
        if (a.phases > 1) {  // phase fastest: the four phases of a pixel tile run on neighbouring CTAs and share its input in L2
          phz = mt & 3;
          mt >>= 2;
          pad_h = 1 - (phz >> 1);
          pad_w = 1 - (phz & 1);
        }
        const int tw = mt % a.tiles_w, th = (mt / a.tiles_w) % a.tiles_h, tn = mt / tiles_hw;
        const int n0 = tn * a.nb, h0 = th * a.bh, w0 = tw * a.bw, cn0 = nt * BN;
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++g) {
          const int s = g % S;
          const uint32_t ph = (g / S) & 1;
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          const uint32_t full = bar_full + 8 * s;
          mbar_expect_tx(full, tx);
          const uint32_t sA = tiles + s * kStageBytes;
          const uint32_t sB = sA + kPlanes * kABytes;
          const int tap = kb / a.cin_blocks, cb = kb - tap * a.cin_blocks;
          const int fy = tap / a.kw, fx = tap - fy * a.kw;
          if (a.stride == 1) {
            const int cx = w0 + fx - pad_w, cy = h0 + fy - pad_h;
            tma_load_4d(sA, &tmAh, full, cb * kBlockK, cx, cy, n0);
            if (SPLIT) tma_load_4d(sA + kABytes, &tmAl, full, cb * kBlockK, cx, cy, n0);
          } else {
            // input row 2*oh + fy - pad = 2*(oh + ay) + py ; same for columns
            const int ty = fy - a.pad_h, tx_ = fx - a.pad_w;
            const int py = ty & 1, px = tx_ & 1;
            const int ay = (ty - py) >> 1, ax = (tx_ - px) >> 1;
            const int cc = px * a.x_cstride + cb * kBlockK;
            tma_load_5d(sA, &tmAh, full, cc, w0 + ax, py, h0 + ay, n0);
            if (SPLIT) tma_load_5d(sA + kABytes, &tmAl, full, cc, w0 + ax, py, h0 + ay, n0);
          }
          if (a.w_per_image | (a.phases > 1)) {
            const int wset = a.w_per_image ? n0 : phz;
            tma_load_3d(sB, &tmBh, full, kb * kBlockK, cn0, wset);
            if (SPLIT) tma_load_3d(sB + kBBytes, &tmBl, full, kb * kBlockK, cn0, wset);
          } else {
            tma_load_2d(sB, &tmBh, full, kb * kBlockK, cn0);
            if (SPLIT) tma_load_2d(sB + kBBytes, &tmBl, full, kb * kBlockK, cn0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (single thread)
    // The K loop of a tile is cut into chunks of a.chunk_kb K-blocks; each chunk accumulates into its own TMEM buffer
    // (the two buffers alternate) and the epilogue warps add the finished chunks in registers.  The tensor core adds into
    // its fp32 accumulator with truncation, so one long accumulation chain loses ~(MMA count) x 2^-25 relative -- 9e-5 for
    // the K = 24048 layers of the 5-frame U-Net, 1.7e-5 at K = 9216; chunks of 1024 K keep every
    // chain at 128-192 MMAs and the cross-chunk sums are round-to-nearest.
    if (lane == 0) {
      uint32_t g = 0;
      int it = 0;  // running chunk counter (accumulator ring position)
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int ks = item % a.ksplit;
        const int kb_lo = ks * a.kb_per_split, kb_hi = min(a.num_kb, kb_lo + a.kb_per_split);
        for (int kb0 = kb_lo; kb0 < kb_hi; kb0 += a.chunk_kb, ++it) {
          const int buf = it & 1;
          mbar_wait(bar_tempty + 8 * buf, ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t tmem_acc = tmem_base + buf * kAccCols;
          const int kb1 = min(kb_hi, kb0 + a.chunk_kb);
          for (int kb = kb0; kb < kb1; ++kb, ++g) {
            const int s = g % S;
            const uint32_t ph = (g / S) & 1;
            mbar_wait(bar_full + 8 * s, ph);
            tc_fence_after();
            const uint32_t sA = tiles + s * kStageBytes;
            const uint32_t sB = sA + kPlanes * kABytes;
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              const uint64_t dAh = umma_desc_sw128(sA + k * 32);
              const uint64_t dBh = umma_desc_sw128(sB + k * 32);
              const uint32_t accum = (kb != kb0 || k != 0) ? 1u : 0u;
              if (kCat) {
                const uint64_t dAl = umma_desc_sw128(sA + kABytes + k * 32);
                umma_f16(tmem_acc, dAh, dBh, kIdescCat, accum);  // N = 2*BN: rows of B_hi then B_lo
                umma_f16(tmem_acc, dAl, dBh, kIdesc, 1);
              } else {
                umma_f16(tmem_acc, dAh, dBh, kIdesc, accum);
                if (SPLIT) {
                  const uint64_t dAl = umma_desc_sw128(sA + kABytes + k * 32);
                  const uint64_t dBl = umma_desc_sw128(sB + kBBytes + k * 32);
                  umma_f16(tmem_acc, dAh, dBl, kIdesc, 1);
                  umma_f16(tmem_acc, dAl, dBh, kIdesc, 1);
                }
              }
            }
            umma_commit(bar_empty + 8 * s);  // frees the smem slot once these MMAs retire
          }
          umma_commit(bar_tfull + 8 * buf);  // chunk accumulator complete
        }
      }
    }
  } else {
    // ===================================================== epilogue: TMEM -> registers -> global
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int ew = warp - 2;  // epilogue warp index; warps ew and ew + 4 share a lane quarter
    const int half = ew >> 2;
    const int r = q * 32 + lane;
    const int wi = r % a.bw, hi_ = (r / a.bw) % a.bh, ni = r / (a.bw * a.bh);
    constexpr int kChunk = BN < 32 ? BN : 32;
    constexpr int kNCH = (BN + 2 * kChunk - 1) / (2 * kChunk);  // 32-column chunks this warp owns: c0 = (2*ci + half) * kChunk
    float4* const stage = a.stage_off < 0 ? nullptr
        : reinterpret_cast<float4*>(smem_raw + (tiles - smem_u32(smem_raw)) + a.stage_off) + ew * 256;
    int it = 0;  // running chunk counter (accumulator ring position)
    int stat_tile = -1;  // tile whose InstanceNorm partial sums sit in s_stat, waiting to be flushed
    constexpr bool split_k = SPLITK;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int tile = item / a.ksplit, ks = item - tile * a.ksplit;
      const int kb_lo = ks * a.kb_per_split, kb_hi = min(a.num_kb, kb_lo + a.kb_per_split);
      const int nt = tile % a.n_tiles;
      int mt = tile / a.n_tiles;
      int oh_off = a.oh_off, ow_off = a.ow_off;
      if (a.phases > 1) {
        const int phz = mt & 3;
        mt >>= 2;
        oh_off = phz >> 1;
        ow_off = phz & 1;
      }
      const int tw = mt % a.tiles_w, th = (mt / a.tiles_w) % a.tiles_h, tn = mt / tiles_hw;
      const int cn0 = nt * BN;
      const int n = tn * a.nb + ni, oh = th * a.bh + hi_, ow = tw * a.bw + wi;
      const bool row_ok = ni < a.nb && n < a.N && oh < a.Ho && ow < a.Wo;
      const long pix = ((long)n * a.out_H + (oh * a.oh_mul + oh_off)) * a.out_W + (ow * a.ow_mul + ow_off);
      // coalesced stores: every lane needs the output offsets of the 8 rows it will write (row = it*4 + lane/8)
      const bool co_ok = kChunk == 32 && stage != nullptr && (a.out_cstride & 3) == 0 &&
                         ((a.out_coffset + cn0) & 3) == 0;  // warp-uniform
      const uint32_t okmask = __ballot_sync(0xffffffffu, row_ok);
      const long mybase = pix * a.out_cstride + a.out_coffset;
      // stage this tile's bias / folded-BN window (all epilogue warps are past the previous tile's reads)
      epi_bar_sync<kEpiWarps * 32>();
      for (int i = threadIdx.x - 64; i < BN; i += kEpiWarps * 32) {
        const int c = cn0 + i;
        const bool ok = c < a.Cout;
        s_bias[i] = (ok && a.bias) ? __ldg(a.bias + c) : 0.f;
        s_scale[i] = (ok && a.scale) ? __ldg(a.scale + c) : 1.f;
        s_shift[i] = (ok && a.scale) ? __ldg(a.shift + c) : 0.f;
        if (a.stats) {  // InstanceNorm statistics: flush the previous finalised tile's partial sums (one image per tile)
          if (stat_tile >= 0) {
            const int ptile = stat_tile;
            const int pc = (ptile % a.n_tiles) * BN + i, pn = (ptile / a.n_tiles) / tiles_hw;
            if (pc < a.Cout) {
              atomicAdd(a.stats + ((long)pn * a.Cout + pc) * 2,
                        (double)s_stat[0][i][0] + (double)s_stat[1][i][0] + (double)s_stat[2][i][0] + (double)s_stat[3][i][0]);
              atomicAdd(a.stats + ((long)pn * a.Cout + pc) * 2 + 1,
                        (double)s_stat[0][i][1] + (double)s_stat[1][i][1] + (double)s_stat[2][i][1] + (double)s_stat[3][i][1]);
            }
          }
        }
      }
      epi_bar_sync<kEpiWarps * 32>();
      stat_tile = -1;
      // ---- all K chunks but the last: partial sums TMEM -> registers (round-to-nearest adds), buffer handed straight back
      // (compiled only into the MULTI instantiations: the single-chunk ones -- every short-mainloop, epilogue-bound layer --
      // keep the register footprint of a plain streaming epilogue)
      float accr[MULTI ? kNCH : 1][32];
      for (int kb0 = kb_lo; MULTI && kb0 + a.chunk_kb < kb_hi; kb0 += a.chunk_kb, ++it) {
        const int buf = it & 1;
        mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + buf * kAccCols;
#pragma unroll
        for (int ci = 0; ci < kNCH; ++ci) {
          const int c0 = (2 * ci + half) * kChunk;
          if (c0 < BN && cn0 + c0 < a.Cout) {  // warp-uniform
            uint32_t v[32];
            const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            if (kChunk == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float acc = (i < kChunk) ? __uint_as_float(v[i]) : 0.f;
              accr[ci][i] = kb0 == kb_lo ? acc : accr[ci][i] + acc;
            }
            if (kCat) {  // right half of the accumulator: hi*lo (the left half holds hi*hi + lo*hi)
              if (kChunk == 32) tmem_ld32(taddr + BN, v); else tmem_ld16(taddr + BN, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < kChunk) accr[ci][i] += __uint_as_float(v[i]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
      }
      // ---- last (or only) chunk: streamed 32 columns at a time straight into bias / activation / folded BN / statistics /
      // stores (both TMEM loads of a column chunk in flight together: short-mainloop layers are bound by this loop)
      const int buf = it & 1;
      mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + buf * kAccCols;
      // split-K: pass 0 parks this item's raw accumulator in its workspace slice, pass 1 (only in the item that arrives last
      // for the tile) re-reads all slices in slice order and runs the epilogue proper; otherwise one pass straight from TMEM
      float* const sk_row = split_k ? a.sk_ws + ((long)mt * kBlockM + r) * a.sk_ctot + cn0 : nullptr;
      bool finalized = !split_k;
#pragma unroll 1
      for (int pass = 0; pass < (split_k ? 2 : 1); ++pass) {
      if (pass == 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);  // accumulator drained: the MMA warp may reuse it
        __threadfence();                                   // this thread's slice stores are visible device-wide ...
        epi_bar_sync<kEpiWarps * 32>();
        if (threadIdx.x == 64) s_flag = atomicAdd(a.sk_cnt + tile, 1);  // ... before the item is counted
        epi_bar_sync<kEpiWarps * 32>();
        if (s_flag != a.ksplit - 1) break;  // CTA-uniform: another item will finalise this tile
        finalized = true;
        __threadfence();
        if (threadIdx.x == 64) a.sk_cnt[tile] = 0;  // ready for the next launch on this workspace
      }
#pragma unroll 1
      for (int ci = 0; ci < kNCH; ++ci) {
        const int c0 = (2 * ci + half) * kChunk;
        if (c0 >= BN || cn0 + c0 >= a.Cout) break;  // warp-uniform
        const int cnt = min(kChunk, a.Cout - (cn0 + c0));  // warp-uniform
        float vals[32];
        if (pass == 0) {
          uint32_t v[32], v2[32];
          const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
          if (kChunk == 32) {
            tmem_ld32(taddr, v);
            if (kCat) tmem_ld32(taddr + BN, v2);
          } else {
            tmem_ld16(taddr, v);
            if (kCat) tmem_ld16(taddr + BN, v2);
          }
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float acc = (i < kChunk) ? __uint_as_float(v[i]) : 0.f;
            if (kCat && i < kChunk) acc += __uint_as_float(v2[i]);  // hi*hi + lo*hi  +  hi*lo
            if (MULTI) {  // earlier chunks' partial sum (register array indexed by the runtime ci through selects)
              float prev = accr[0][i];
#pragma unroll
              for (int k = 1; k < (MULTI ? kNCH : 1); ++k) prev = ci == k ? accr[k][i] : prev;
              acc += prev;
            }
            vals[i] = acc;
          }
          if (split_k) {  // kChunk is a multiple of 16: float4 stores of this lane's row segment
            float4* dst = reinterpret_cast<float4*>(sk_row + (long)ks * a.sk_slice + c0);
#pragma unroll
            for (int i = 0; i < kChunk; i += 4) __stcg(dst + i / 4, make_float4(vals[i], vals[i + 1], vals[i + 2], vals[i + 3]));
            continue;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) vals[i] = 0.f;
          for (int k2 = 0; k2 < a.ksplit; k2 += 2) {  // two slices' loads in flight, added in slice order
            const float4* src = reinterpret_cast<const float4*>(sk_row + (long)k2 * a.sk_slice + c0);
            const bool two = k2 + 1 < a.ksplit;
            const float4* src2 = two ? src + a.sk_slice / 4 : src;
            float4 ta[kChunk / 4], tb[kChunk / 4];
#pragma unroll
            for (int i = 0; i < kChunk / 4; ++i) ta[i] = __ldcg(src + i);
#pragma unroll
            for (int i = 0; i < kChunk / 4; ++i) tb[i] = __ldcg(src2 + i);
#pragma unroll
            for (int i = 0; i < kChunk / 4; ++i) {
              vals[4 * i] += ta[i].x; vals[4 * i + 1] += ta[i].y; vals[4 * i + 2] += ta[i].z; vals[4 * i + 3] += ta[i].w;
            }
            if (two) {
#pragma unroll
              for (int i = 0; i < kChunk / 4; ++i) {
                vals[4 * i] += tb[i].x; vals[4 * i + 1] += tb[i].y; vals[4 * i + 2] += tb[i].z; vals[4 * i + 3] += tb[i].w;
              }
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i)
          vals[i] = (i < kChunk) ? fmaf(vals[i], a.acc_scale, s_bias[c0 + (i < kChunk ? i : 0)]) : 0.f;
        act_chunk_dispatch(vals, a.pre_act, a.act_param);
        if (a.scale != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < kChunk) vals[i] = fmaf(vals[i], s_scale[c0 + i], s_shift[c0 + i]);
        }
        act_chunk_dispatch(vals, a.post_act, a.act_param);
        if (a.stats != nullptr) {  // warp-uniform: column sums of this 32-row x 32-channel chunk -> shared partials
          float sv[32];  // one scratch array, used twice (register budget: 168 per thread with 10 warps)
#pragma unroll
          for (int i = 0; i < 32; ++i) sv[i] = (row_ok && i < cnt) ? vals[i] : 0.f;
          s_stat[q][c0 + lane][0] = warp_transpose_sum(sv, lane);  // single writer of (quarter q, chunk c0); lanes >= cnt: 0
#pragma unroll
          for (int i = 0; i < 32; ++i) sv[i] = (row_ok && i < cnt) ? vals[i] * vals[i] : 0.f;
          s_stat[q][c0 + lane][1] = warp_transpose_sum(sv, lane);
        }
        if (co_ok && (cnt & 3) == 0)
          conv_store_chunk_coalesced(vals, stage, lane, mybase, okmask, cn0 + c0, cnt, a);
        else if (row_ok)
          conv_store_row(vals, cn0 + c0, cnt, pix, a);
      }
      }
      if (a.stats != nullptr && finalized) stat_tile = tile;  // CTA-uniform
      if (!split_k) {  // hand the accumulator back to the MMA warp (split-K did so before counting the item)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
      }
      ++it;
    }
    if (a.stats != nullptr) epi_bar_sync<kEpiWarps * 32>();
    if (a.stats != nullptr && stat_tile >= 0) {  // the last finalised tile's statistics
      const int ptile = stat_tile;
      for (int i = threadIdx.x - 64; i < BN; i += kEpiWarps * 32) {
        const int pc = (ptile % a.n_tiles) * BN + i, pn = (ptile / a.n_tiles) / tiles_hw;
        if (pc < a.Cout) {
          atomicAdd(a.stats + ((long)pn * a.Cout + pc) * 2,
                    (double)s_stat[0][i][0] + (double)s_stat[1][i][0] + (double)s_stat[2][i][0] + (double)s_stat[3][i][0]);
          atomicAdd(a.stats + ((long)pn * a.Cout + pc) * 2 + 1,
                    (double)s_stat[0][i][1] + (double)s_stat[1][i][1] + (double)s_stat[2][i][1] + (double)s_stat[3][i][1]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// First-layer variant: the A operand is produced IN the kernel.  Small-Cin layers (3/10/22 input channels, FlowNet
// stems) are a dense GEMM over K = kh*kw*Cin; instead of materialising the im2col matrix in HBM (1.5 GB for the
// GMM's 22-channel input at 80 frames) four producer warps gather the NCHW f32 input, split it to 16-bit hi/lo
// and write the 128x64 K-major tile straight into shared memory in the SWIZZLE_128B pattern tcgen05.mma expects
// (16-byte chunk j of row r lives at chunk j ^ (r & 7)); B still arrives by TMA.
// Warp roles: 0-7 A producers (two threads per tile row, 32 k each), 8 = B TMA, 9 = TMEM alloc + MMA issuer,
// 10-13 = epilogue.
// ---------------------------------------------------------------------------------------------------------
constexpr int kI2cProdWarps = 8;  // 2 threads per tile row: enough loads in flight to cover the L2/HBM latency
constexpr int kI2cThreads = (kI2cProdWarps + 6) * 32;

struct Im2colSrc {
  const float* x0;
  const float* x1;
  int C0, C1, H, W;  // input NCHW geometry (two tensors concatenated on C)
};

template <int BN, bool SPLIT>
__global__ void __launch_bounds__(kI2cThreads, 1)
    conv_im2col_igemm_kernel(const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                             const ConvArgs a, const Im2colSrc src) {
  pdl_grid_sync();
  constexpr int kBBytes = BN * kBlockK * 2;
  constexpr int kPlanes = SPLIT ? 2 : 1;
  constexpr int kStageBytes = kPlanes * (kABytes + kBBytes);
  constexpr int kAccCols = BN < 32 ? 32 : BN;
  constexpr int kTmemCols = 2 * kAccCols;
  const uint32_t kIdesc = a.idesc;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 4];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[BN], s_scale[BN], s_shift[BN];

  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles_ptr = smem_raw + (tiles - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.stages;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[kMaxStages]),
                 bar_tfull = smem_u32(&bars[2 * kMaxStages]), bar_tempty = smem_u32(&bars[2 * kMaxStages + 2]);

  if (warp == kI2cProdWarps && lane == 0) {
    prefetch_tmap(&tmBh);
    if (SPLIT) prefetch_tmap(&tmBl);
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, kI2cProdWarps + 1);  // producer warps + the TMA thread's arrive.expect_tx
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kI2cProdWarps + 1) tmem_alloc<kTmemCols>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int tiles_hw = a.tiles_w * a.tiles_h;
  const int C = src.C0 + src.C1;
  const int K = a.kh * a.kw * C;
  const int HW = src.H * src.W;

  if (warp < kI2cProdWarps) {
    // ===================================================== A producers: gather + split + swizzled store
    const int r = threadIdx.x & 127;    // tile row
    const int half = threadIdx.x >> 7;  // which 32-wide half of the 64-wide K block this thread fills
    const int wi = r % a.bw, hi_ = (r / a.bw) % a.bh, ni = r / (a.bw * a.bh);
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const int mt = tile / a.n_tiles;
      const int tw = mt % a.tiles_w, th = (mt / a.tiles_w) % a.tiles_h, tn = mt / tiles_hw;
      const int n = tn * a.nb + ni, oh = th * a.bh + hi_, ow = tw * a.bw + wi;
      const bool row_ok = ni < a.nb && n < a.N && oh < a.Ho && ow < a.Wo;
      const int iy0 = oh * a.stride - a.pad_h, ix0 = ow * a.stride - a.pad_w;
      const float* p0 = src.x0 + (long)n * src.C0 * HW;
      const float* p1 = src.x1 ? src.x1 + (long)n * src.C1 * HW : src.x0;
      for (int kb = 0; kb < a.num_kb; ++kb, ++g) {
        const int s = g % S;
        const uint32_t ph = (g / S) & 1;
        int k = kb * kBlockK + half * 32;
        int tap = k / C, c = k - tap * C;
        int fy = tap / a.kw, fx = tap - fy * a.kw;
        // issue all 32 gathers of this thread before anything consumes them
        float v[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          float val = 0.f;
          const int iy = iy0 + fy, ix = ix0 + fx;
          if (row_ok && k + e < K && iy >= 0 && iy < src.H && ix >= 0 && ix < src.W)
            val = c < src.C0 ? __ldg(p0 + (long)c * HW + iy * src.W + ix) : __ldg(p1 + (long)(c - src.C0) * HW + iy * src.W + ix);
          v[e] = val;
          if (++c == C) {
            c = 0;
            if (++fx == a.kw) { fx = 0; ++fy; }
          }
        }
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        uint8_t* rowh = tiles_ptr + s * kStageBytes + r * 128;
        uint8_t* rowl = rowh + kABytes;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          __align__(16) plane_t hi[8];
          __align__(16) plane_t lo[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) split16(v[jj * 8 + e], a.fmt, hi[e], lo[e]);
          const int chunk = ((half * 4 + jj) ^ (r & 7)) * 16;
          *reinterpret_cast<uint4*>(rowh + chunk) = *reinterpret_cast<const uint4*>(hi);
          if (SPLIT) *reinterpret_cast<uint4*>(rowl + chunk) = *reinterpret_cast<const uint4*>(lo);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * s);
      }
    }
  } else if (warp == kI2cProdWarps) {
    // ===================================================== B operand by TMA
    if (lane == 0) {
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const int cn0 = (tile % a.n_tiles) * BN;
        for (int kb = 0; kb < a.num_kb; ++kb, ++g) {
          const int s = g % S;
          const uint32_t ph = (g / S) & 1;
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          const uint32_t full = bar_full + 8 * s;
          mbar_expect_tx(full, kPlanes * kBBytes);
          const uint32_t sB = tiles + s * kStageBytes + kPlanes * kABytes;
          tma_load_2d(sB, &tmBh, full, kb * kBlockK, cn0);
          if (SPLIT) tma_load_2d(sB + kBBytes, &tmBl, full, kb * kBlockK, cn0);
        }
      }
    }
  } else if (warp == kI2cProdWarps + 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      uint32_t g = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(bar_tempty + 8 * buf, ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + buf * kAccCols;
        for (int kb = 0; kb < a.num_kb; ++kb, ++g) {
          const int s = g % S;
          const uint32_t ph = (g / S) & 1;
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          const uint32_t sA = tiles + s * kStageBytes;
          const uint32_t sB = sA + kPlanes * kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t dAh = umma_desc_sw128(sA + k * 32);
            const uint64_t dBh = umma_desc_sw128(sB + k * 32);
            umma_f16(tmem_acc, dAh, dBh, kIdesc, (kb | k) != 0);
            if (SPLIT) {
              const uint64_t dAl = umma_desc_sw128(sA + kABytes + k * 32);
              const uint64_t dBl = umma_desc_sw128(sB + kBBytes + k * 32);
              umma_f16(tmem_acc, dAh, dBl, kIdesc, 1);
              umma_f16(tmem_acc, dAl, dBh, kIdesc, 1);
            }
          }
          umma_commit(bar_empty + 8 * s);
        }
        umma_commit(bar_tfull + 8 * buf);
      }
    }
  } else {
    // ===================================================== epilogue (same as conv_igemm_kernel)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int wi = r % a.bw, hi_ = (r / a.bw) % a.bh, ni = r / (a.bw * a.bh);
    constexpr int kChunk = BN < 32 ? BN : 32;
    const int et = threadIdx.x - (kI2cProdWarps + 2) * 32;  // 0..127
    int it = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
      const int nt = tile % a.n_tiles, mt = tile / a.n_tiles;
      const int tw = mt % a.tiles_w, th = (mt / a.tiles_w) % a.tiles_h, tn = mt / tiles_hw;
      const int cn0 = nt * BN;
      const int n = tn * a.nb + ni, oh = th * a.bh + hi_, ow = tw * a.bw + wi;
      const bool row_ok = ni < a.nb && n < a.N && oh < a.Ho && ow < a.Wo;
      const long pix = ((long)n * a.out_H + (oh * a.oh_mul + a.oh_off)) * a.out_W + (ow * a.ow_mul + a.ow_off);
      epi_bar_sync<128>();
      for (int i = et; i < BN; i += 128) {
        const int c = cn0 + i;
        const bool ok = c < a.Cout;
        s_bias[i] = (ok && a.bias) ? __ldg(a.bias + c) : 0.f;
        s_scale[i] = (ok && a.scale) ? __ldg(a.scale + c) : 1.f;
        s_shift[i] = (ok && a.scale) ? __ldg(a.shift + c) : 0.f;
      }
      epi_bar_sync<128>();
      const int buf = it & 1;
      mbar_wait(bar_tfull + 8 * buf, (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + buf * kAccCols;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += kChunk) {
        if (cn0 + c0 >= a.Cout) break;
        uint32_t v[32];
        const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        if (kChunk == 32)
          tmem_ld32(taddr, v);
        else
          tmem_ld16(taddr, v);
        tmem_ld_wait();
        const int cnt = min(kChunk, a.Cout - (cn0 + c0));
        float vals[32];
#pragma unroll
        for (int i = 0; i < 32; ++i)
          vals[i] = (i < kChunk) ? fmaf(__uint_as_float(v[i]), a.acc_scale, s_bias[c0 + (i < kChunk ? i : 0)]) : 0.f;
        act_chunk_dispatch(vals, a.pre_act, a.act_param);
        if (a.scale != nullptr) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < kChunk) vals[i] = fmaf(vals[i], s_scale[c0 + i], s_shift[c0 + i]);
        }
        act_chunk_dispatch(vals, a.post_act, a.act_param);
        if (row_ok) conv_store_row(vals, cn0 + c0, cnt, pix, a);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kI2cProdWarps + 1) tmem_dealloc<kTmemCols>(tmem_base);
}

// ---------------------------------------------------------------------------- CUDA-core cross-check
// Same operands, same epilogue, plain fp32 FMAs.  Tests compare the tcgen05 kernel against this on
// the GPU at sizes where the CPU oracle would take minutes.  Not on the product path.
__global__ void __launch_bounds__(128)
    conv_direct_kernel(const plane_t* __restrict__ xh, const plane_t* __restrict__ xl,
                       const plane_t* __restrict__ wh, const plane_t* __restrict__ wl, int H, int W,
                       ConvArgs a) {
  pdl_grid_sync();
  const long total = (long)a.N * a.Ho * a.Wo * a.Cout;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int c = (int)(e % a.Cout);
    const long p = e / a.Cout;
    const int ow = (int)(p % a.Wo), oh = (int)((p / a.Wo) % a.Ho), n = (int)(p / ((long)a.Wo * a.Ho));
    float acc = 0.f;
    for (int fy = 0; fy < a.kh; ++fy) {
      const int iy = oh * a.stride + fy - a.pad_h;
      if (iy < 0 || iy >= H) continue;
      for (int fx = 0; fx < a.kw; ++fx) {
        const int ix = ow * a.stride + fx - a.pad_w;
        if (ix < 0 || ix >= W) continue;
        const long xo = (((long)n * H + iy) * W + ix) * a.x_cstride;
        const long wo = (((long)(a.w_per_image ? n : 0) * a.Cout + c) * a.kh * a.kw + fy * a.kw + fx) * a.cin_pad;
        for (int ci = 0; ci < a.cin_pad; ++ci) {
          const float xhv = load16(xh[xo + ci], a.fmt), whv = load16(wh[wo + ci], a.fmt);
          acc = fmaf(xhv, whv, acc);
          if (xl) {
            acc = fmaf(xhv, load16(wl[wo + ci], a.fmt), acc);
            acc = fmaf(load16(xl[xo + ci], a.fmt), whv, acc);
          }
        }
      }
    }
    float v = conv_epilogue_value(acc, c, a);
    const long pix = ((long)n * a.out_H + (oh * a.oh_mul + a.oh_off)) * a.out_W + (ow * a.ow_mul + a.ow_off);
    const long o = pix * a.out_cstride + a.out_coffset + c;
    if (a.y_f32) a.y_f32[o] = v;
    if (a.y_hi) {
      plane_t h, l;
      split16(v, a.fmt, h, l);
      a.y_hi[o] = h;
      if (a.y_lo) a.y_lo[o] = l;
    }
  }
}

// ------------------------------------------------------------------------------- weight packing
// One thread per (output channel, packed input channel): its kh*kw taps are consecutive floats of the source (a warp reads
// 32 neighbouring runs: every fetched sector is used across the tap loop) and each tap's store is coalesced over the packed
// channels.  (The first version walked the destination linearly with three 64-bit divisions per element and a
// kh*kw-strided gather: 0.43 ms per training step for 45 M elements, profiles/r02_train_launches_summary.md.)
__global__ void __launch_bounds__(128)
    pack_conv_weight_kernel(const float* __restrict__ w, plane_t* __restrict__ whi,
                            plane_t* __restrict__ wlo, int Cout, int Cin, int kh, int kw, int cin_pad,
                            const int32_t* __restrict__ chan_map, int transpose_io, int fmt, float w_scale) {
  pdl_grid_sync();
  const int taps = kh * kw;
  for (int co = blockIdx.y; co < Cout; co += gridDim.y) {
    for (int cp = blockIdx.x * blockDim.x + threadIdx.x; cp < cin_pad; cp += gridDim.x * blockDim.x) {
      const int ci = chan_map ? chan_map[cp] : (cp < Cin ? cp : -1);
      const bool ok = ci >= 0 && ci < Cin;
      // ConvTranspose2d weight [Cin][Cout][kh][kw] is read with its taps flipped
      const float* src = ok ? (transpose_io ? w + ((long)ci * Cout + co) * taps : w + ((long)co * Cin + ci) * taps) : w;
      const long dst = (long)co * taps * cin_pad + cp;
      for (int t = 0; t < taps; ++t) {
        const float v = ok ? __ldg(src + (transpose_io ? taps - 1 - t : t)) : 0.f;
        plane_t h, l;
        split16(v * w_scale, fmt, h, l);
        whi[dst + (long)t * cin_pad] = h;
        if (wlo) wlo[dst + (long)t * cin_pad] = l;
      }
    }
  }
}

// ConvTranspose2d(4, 2, 1) weight [Cin][Cout][4][4] -> the four phase-wise 2x2 stride-1 convs it decomposes into
// (networks/deconv.py): out[phase = py*2+px][co][(fy,fx)][ci_pad] = w[ci][co][T[py][fy]][T[px][fx]], T = {0:{3,1}, 1:{2,0}}.
__global__ void __launch_bounds__(256)
    pack_deconv4x4s2_weight_kernel(const float* __restrict__ w, plane_t* __restrict__ whi, plane_t* __restrict__ wlo,
                                   int Cout, int Cin, int cin_pad, const int32_t* __restrict__ chan_map, int fmt,
                                   float w_scale) {
  pdl_grid_sync();
  const long per_phase = (long)Cout * 4 * cin_pad;
  const long total = 4 * per_phase;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int cp = (int)(e % cin_pad);
    const int tap = (int)((e / cin_pad) & 3);
    const int co = (int)((e / ((long)cin_pad * 4)) % Cout);
    const int phase = (int)(e / per_phase);
    const int py = phase >> 1, px = phase & 1, fy = tap >> 1, fx = tap & 1;
    const int ty = py ? (fy ? 0 : 2) : (fy ? 1 : 3);
    const int tx = px ? (fx ? 0 : 2) : (fx ? 1 : 3);
    const int ci = chan_map ? chan_map[cp] : (cp < Cin ? cp : -1);
    float v = 0.f;
    if (ci >= 0 && ci < Cin) v = w[(((long)ci * Cout + co) * 4 + ty) * 4 + tx];
    plane_t h, l;
    split16(v * w_scale, fmt, h, l);
    whi[e] = h;
    if (wlo) wlo[e] = l;
  }
}

constexpr int kSplitKMinKb = 32;  // K-blocks below which the K loop is too short to be worth a second pass
constexpr int kSplitKMax = 16;    // slices per tile (the finalising CTA reads them all)

static int sm_count() {
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
  }
  return num_sms;
}

// Split-K plan: layers whose tiles fill less than half of the SMs and whose K loop is deep (the 4x3 ... 8x6 levels of the
// U-Net / GMM regression / FlowNet pyramids: 8-40 tiles, 36-150 K-blocks each) cut K into slices of >= 8 K-blocks so that
// tiles x slices covers the chip.  Returns the slice count (1 = off) and the K-blocks per slice.
static int plan_splitk(int total_tiles, int num_kb, int bn, int chunk_kb, int* kb_per_split) {
  *kb_per_split = num_kb;
  const int sms = sm_count();
  if ((bn != 64 && bn != 128) || total_tiles * 2 > sms || num_kb < kSplitKMinKb) return 1;
  int ks = sms / total_tiles;
  if (ks > num_kb / 8) ks = num_kb / 8;
  if (ks > kSplitKMax) ks = kSplitKMax;
  if (ks < 2) return 1;
  int per = cdiv(num_kb, ks);
  if (per > chunk_kb) per = chunk_kb;  // one TMEM accumulation chain per slice (the split-K kernels are not MULTI)
  *kb_per_split = per;
  return cdiv(num_kb, per);
}
static size_t splitk_bytes(int total_tiles, int m_tiles, int n_tiles, int bn, int ksplit) {
  const size_t cnt = ((size_t)total_tiles * sizeof(int) + 255) & ~(size_t)255;
  return cnt + (size_t)ksplit * m_tiles * kBlockM * ((size_t)n_tiles * bn) * sizeof(float);
}

// Picks the (nb, bh, bw) pixel tile with the fewest wasted accumulator rows.
static void pick_tile(int N, int Ho, int Wo, int& nb, int& bh, int& bw) {
  double best = -1.0;
  nb = 1; bh = 1; bw = 1;
  for (int w = 1; w <= Wo && w <= kBlockM; ++w) {
    for (int h = 1; h <= Ho && h * w <= kBlockM; ++h) {
      int n = kBlockM / (w * h);
      if (n > N) n = N;
      if (n < 1) n = 1;
      long tiles = (long)cdiv(Wo, w) * cdiv(Ho, h) * cdiv(N, n);
      double eff = (double)N * Ho * Wo / ((double)tiles * kBlockM);
      // prefer wide tiles on ties (longer contiguous runs for TMA)
      eff += 1e-6 * w;
      if (eff > best) { best = eff; nb = n; bh = h; bw = w; }
    }
  }
}

template <int BN, bool SPLIT, bool MULTI, bool SPLITK = false>
static int launch_igemm(const CUtensorMap& tAh, const CUtensorMap& tAl, const CUtensorMap& tBh, const CUtensorMap& tBl,
                        ConvArgs& a, int m_tiles, int stages_req, cudaStream_t stream) {
  constexpr int kBBytes = BN * kBlockK * 2;
  constexpr int kStageBytes = (SPLIT ? 2 : 1) * (kABytes + kBBytes);
  int stages = (200 * 1024) / kStageBytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages_req > 0 && stages_req < stages) stages = stages_req;
  if (stages > a.num_kb) stages = a.num_kb;  // (more stages than K-blocks on short-K layers: measured, no gain, r02)
  if (stages < 1) stages = 1;
  a.stages = stages;
  a.idesc = umma_idesc_f16(kBlockM, BN, a.fmt == SHINEON_FMT_FP16 ? 0 : 1);
  a.idesc_cat = umma_idesc_f16(kBlockM, 2 * BN <= 256 ? 2 * BN : BN, a.fmt == SHINEON_FMT_FP16 ? 0 : 1);
  static int max_dyn_smem = -1;  // per instantiation: opt-in limit minus this kernel's static shared memory
  if (max_dyn_smem < 0) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, conv_igemm_kernel<BN, SPLIT, MULTI, SPLITK>);
    if (e == cudaSuccess) {
      max_dyn_smem = 227 * 1024 - (int)fa.sharedSizeBytes;
      e = cudaFuncSetAttribute(conv_igemm_kernel<BN, SPLIT, MULTI, SPLITK>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn_smem);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      max_dyn_smem = -1;
      return fail(SHINEON_ERR_CUDA, "conv_igemm shared-memory opt-in: %s", cudaGetErrorString(e));
    }
  }
  while (stages > 1 && stages * kStageBytes + 1024 > max_dyn_smem) --stages;
  // f32-only outputs of short-mainloop layers are epilogue-bound: give them the transpose buffers for coalesced
  // stores (4 KB per epilogue warp) when that leaves at least min(num_kb, 2) pipeline stages.  (Plane outputs through the
  // same buffers -- 8 lanes x 8 bytes per pixel and plane -- were measured in r02: the try-on step got 2 % slower.)
  constexpr int kStagingBytes = kEpiWarps * 4096;
  a.stage_off = -1;
  if (a.y_hi == nullptr && a.y_f32 != nullptr && BN >= 32 && a.num_kb <= 16) {
    int st2 = stages;
    while (st2 > 1 && st2 * kStageBytes + 1024 + kStagingBytes > max_dyn_smem) --st2;
    if (st2 * kStageBytes + 1024 + kStagingBytes <= max_dyn_smem && st2 >= (a.num_kb < 2 ? a.num_kb : 2)) {
      stages = st2;
      a.stage_off = stages * kStageBytes;
    }
  }
  a.stages = stages;
  const int smem_bytes = stages * kStageBytes + 1024 + (a.stage_off >= 0 ? kStagingBytes : 0);
  if (smem_bytes > max_dyn_smem) return fail(SHINEON_ERR_UNSUPPORTED, "conv_igemm: tile does not fit in shared memory");
  a.n_tiles = cdiv(a.Cout, BN);
  const long total = (long)m_tiles * a.n_tiles;
  if (total >= (1l << 31)) return fail(SHINEON_ERR_ARG, "conv2d: too many tiles");
  a.total_tiles = (int)total;
  const int num_sms = sm_count();
  const long items = (long)a.total_tiles * a.ksplit;
  const int grid = items < num_sms ? (int)items : num_sms;
  klaunch(conv_igemm_kernel<BN, SPLIT, MULTI, SPLITK>, grid, kThreads, smem_bytes, stream, tAh, tAl, tBh, tBl, a);
  return after_launch("conv_igemm_kernel");
}

static int fill_args(const shineon_conv2d_params* p, ConvArgs& a) {
  SHINEON_REQUIRE(p != nullptr, "conv2d: null params");
  SHINEON_REQUIRE(p->x_hi && p->w_hi, "conv2d: null operand");
  SHINEON_REQUIRE((p->x_lo == nullptr) == (p->w_lo == nullptr), "conv2d: x_lo and w_lo must both be given or both NULL");
  SHINEON_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0 && p->Cout > 0, "conv2d: bad shape");
  SHINEON_REQUIRE(p->cin_pad > 0 && p->cin_pad % kBlockK == 0, "conv2d: cin_pad %d must be a multiple of 64", p->cin_pad);
  SHINEON_REQUIRE(p->kh > 0 && p->kw > 0 && p->kh <= 7 && p->kw <= 7, "conv2d: kernel %dx%d unsupported", p->kh, p->kw);
  SHINEON_REQUIRE(p->stride == 1 || p->stride == 2, "conv2d: stride %d unsupported", p->stride);
  SHINEON_REQUIRE(p->pad_h >= 0 && p->pad_w >= 0 && p->pad_h < p->kh + p->stride && p->pad_w < p->kw + p->stride, "conv2d: pad");
  SHINEON_REQUIRE(p->Ho > 0 && p->Wo > 0 && (p->Ho - 1) * p->stride - p->pad_h < p->H && (p->Wo - 1) * p->stride - p->pad_w < p->W,
                  "conv2d: Ho/Wo (%d,%d) reach past the input", p->Ho, p->Wo);
  SHINEON_REQUIRE(p->stride == 1 || (p->H % 2 == 0 && p->W % 2 == 0), "conv2d: stride 2 needs even H,W");
  SHINEON_REQUIRE(p->y_f32 || p->y_hi, "conv2d: no output");
  SHINEON_REQUIRE((p->scale == nullptr) == (p->shift == nullptr), "conv2d: scale/shift");
  SHINEON_REQUIRE((reinterpret_cast<uintptr_t>(p->x_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->w_hi) & 15) == 0, "conv2d: operands must be 16-byte aligned");
  a.N = p->N; a.Ho = p->Ho; a.Wo = p->Wo; a.Cout = p->Cout;
  a.kh = p->kh; a.kw = p->kw; a.stride = p->stride; a.pad_h = p->pad_h; a.pad_w = p->pad_w;
  a.cin_pad = p->cin_pad; a.cin_blocks = p->cin_pad / kBlockK;
  a.x_cstride = p->x_cstride ? p->x_cstride : p->cin_pad;
  SHINEON_REQUIRE(a.x_cstride >= a.cin_pad && a.x_cstride % 8 == 0, "conv2d: x_cstride %d", a.x_cstride);
  a.num_kb = p->kh * p->kw * a.cin_blocks;
  a.chunk_kb = p->acc_chunk_kb > 0 ? p->acc_chunk_kb : (p->acc_chunk_kb < 0 ? a.num_kb : 16);
  a.stages = 1;
  a.bias = p->bias; a.scale = p->scale; a.shift = p->shift;
  a.pre_act = p->pre_act; a.post_act = p->post_act; a.act_param = p->act_param;
  SHINEON_REQUIRE(p->plane_fmt == SHINEON_FMT_BF16 || p->plane_fmt == SHINEON_FMT_FP16, "conv2d: plane_fmt %d", p->plane_fmt);
  a.fmt = p->plane_fmt;
  a.acc_scale = p->acc_scale == 0.f ? 1.f : p->acc_scale;
  a.idesc = 0;
  a.idesc_cat = 0;
  a.stage_off = -1;
  a.y_f32 = p->y_f32; a.y_hi = (plane_t*)p->y_hi; a.y_lo = (plane_t*)p->y_lo;
  a.oh_mul = p->oh_mul ? p->oh_mul : 1; a.ow_mul = p->ow_mul ? p->ow_mul : 1;
  a.oh_off = p->oh_off; a.ow_off = p->ow_off;
  a.out_H = p->out_H ? p->out_H : p->Ho; a.out_W = p->out_W ? p->out_W : p->Wo;
  a.out_cstride = p->out_cstride ? p->out_cstride : p->Cout;
  a.out_coffset = p->out_coffset;
  SHINEON_REQUIRE(a.out_coffset >= 0 && a.out_coffset + a.Cout <= a.out_cstride, "conv2d: output channel window out of range");
  SHINEON_REQUIRE(p->deconv_phases || ((a.Ho - 1) * a.oh_mul + a.oh_off < a.out_H && (a.Wo - 1) * a.ow_mul + a.ow_off < a.out_W),
                  "conv2d: output pixel window out of range");
  a.w_per_image = p->w_per_image ? 1 : 0;
  a.phases = p->deconv_phases ? 4 : 1;
  if (a.phases > 1) {
    SHINEON_REQUIRE(p->kh == 2 && p->kw == 2 && p->stride == 1 && p->Ho == p->H && p->Wo == p->W && !p->w_per_image,
                    "conv2d: deconv_phases needs the 2x2 / stride 1 phase geometry (Ho = H, Wo = W)");
    SHINEON_REQUIRE(a.out_H == 2 * p->H && a.out_W == 2 * p->W && p->stats_ws == nullptr,
                    "conv2d: deconv_phases writes a [N, 2H, 2W] output and has no statistics epilogue");
    a.oh_mul = 2; a.ow_mul = 2; a.oh_off = 0; a.ow_off = 0;  // the phase of a tile sets the offsets
    a.pad_h = 1; a.pad_w = 1;
  }
  pick_tile(a.w_per_image ? 1 : a.N, a.Ho, a.Wo, a.nb, a.bh, a.bw);
  a.stats = nullptr;
  a.tiles_w = cdiv(a.Wo, a.bw);
  a.tiles_h = cdiv(a.Ho, a.bh);
  a.phase_tiles = a.tiles_w * a.tiles_h * cdiv(a.N, a.nb);
  return SHINEON_OK;
}

}  // namespace shineon

using namespace shineon;

int launch_instnorm_stats(const float* x, double* ws, int N, int HW, int C, cudaStream_t stream);  // norm_act.cu

static int conv2d_igemm_launch(const shineon_conv2d_params* p, ConvArgs& a, cudaStream_t stream);

extern "C" int shineon_conv2d_igemm_fwd(const shineon_conv2d_params* p, shineon_stream_t stream_) {
  ConvArgs a;
  int rc = fill_args(p, a);
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  // InstanceNorm statistics of the f32 output: in the epilogue when every pixel tile lies inside one image, otherwise
  // (the 4x3 ... 8x6 levels, a few hundred KB) by the standalone pass right behind the conv
  bool stats_after = false;
  if (p->stats_ws != nullptr) {
    if (a.nb == 1) {
      a.stats = p->stats_ws;
    } else {
      SHINEON_REQUIRE(p->y_f32 && a.out_cstride == a.Cout && a.out_H == a.Ho && a.out_W == a.Wo,
                      "conv2d: stats_ws on a multi-image tile needs a dense f32 output");
      stats_after = true;
    }
  }
  rc = conv2d_igemm_launch(p, a, stream);
  if (rc == SHINEON_OK && stats_after) rc = launch_instnorm_stats(p->y_f32, p->stats_ws, a.N, a.Ho * a.Wo, a.Cout, stream);
  return rc;
}

extern "C" size_t shineon_conv2d_splitk_workspace_bytes(const shineon_conv2d_params* p) {
  ConvArgs a;
  if (fill_args(p, a) != SHINEON_OK || a.phases > 1) return 0;
  int bn = p->tile_n;
  if (bn == 0) bn = a.Cout <= 16 ? 16 : a.Cout <= 32 ? 32 : a.Cout <= 64 ? 64 : 128;
  const long m_tiles = (long)a.tiles_w * a.tiles_h * cdiv(a.N, a.nb);
  const int n_tiles = cdiv(a.Cout, bn);
  if (m_tiles * n_tiles > 4096) return 0;
  int per;
  const int ks = plan_splitk((int)(m_tiles * n_tiles), a.num_kb, bn, a.chunk_kb, &per);
  return ks > 1 ? splitk_bytes((int)(m_tiles * n_tiles), (int)m_tiles, n_tiles, bn, ks) : 0;
}

static int conv2d_igemm_launch(const shineon_conv2d_params* p, ConvArgs& a, cudaStream_t stream) {
  int rc;
  const bool split = p->x_lo != nullptr;
  SHINEON_REQUIRE((long)a.phases * a.phase_tiles < (1l << 31), "conv2d: too many tiles");
  const int m_tiles = a.phases * a.phase_tiles;

  int bn = p->tile_n;
  if (bn == 0) bn = a.Cout <= 16 ? 16 : a.Cout <= 32 ? 32 : a.Cout <= 64 ? 64 : 128;
  SHINEON_REQUIRE(bn == 16 || bn == 32 || bn == 64 || bn == 128 || bn == 256, "conv2d: tile_n %d", bn);

  // ---- tensor maps
  CUtensorMap tAh, tAl, tBh, tBl;
  // C = channel extent the boxes may touch, Cs = pixel pitch (both in elements)
  const cuuint64_t Cw = (cuuint64_t)p->cin_pad, C = (cuuint64_t)a.x_cstride, H = (cuuint64_t)p->H, W = (cuuint64_t)p->W, N = (cuuint64_t)p->N;
  // a channel window of a wider (concat) buffer: the bytes next to a pixel's window belong to another tensor, so a 256-byte L2
  // promotion would fetch them for nothing (measured: 2.0x DRAM over-read on the 64-of-128-channel U-Net skip window)
  const bool promo = a.x_cstride == p->cin_pad;
  if (p->stride == 1) {
    cuuint64_t dims[4] = {Cw, W, H, N};
    cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)a.bw, (cuuint32_t)a.bh, (cuuint32_t)a.nb};
    if ((rc = encode_map(&tAh, p->x_hi, 4, dims, strides, box, "A hi", p->plane_fmt, promo))) return rc;
    if (split && (rc = encode_map(&tAl, p->x_lo, 4, dims, strides, box, "A lo", p->plane_fmt, promo))) return rc;
  } else {
    cuuint64_t dims[5] = {C + Cw, W / 2, 2, H / 2, N};  // column pair: [q*Cs + c], c < cin_pad
    cuuint64_t strides[4] = {2 * C * 2, W * C * 2, 2 * W * C * 2, H * W * C * 2};
    cuuint32_t box[5] = {(cuuint32_t)kBlockK, (cuuint32_t)a.bw, 1, (cuuint32_t)a.bh, (cuuint32_t)a.nb};
    if ((rc = encode_map(&tAh, p->x_hi, 5, dims, strides, box, "A hi s2", p->plane_fmt, promo))) return rc;
    if (split && (rc = encode_map(&tAl, p->x_lo, 5, dims, strides, box, "A lo s2", p->plane_fmt, promo))) return rc;
  }
  {
    const cuuint64_t K = (cuuint64_t)p->kh * p->kw * Cw;
    const int brank = (a.w_per_image || a.phases > 1) ? 3 : 2;
    cuuint64_t dims[3] = {K, (cuuint64_t)p->Cout, a.phases > 1 ? (cuuint64_t)a.phases : N};
    cuuint64_t strides[2] = {K * 2, K * 2 * (cuuint64_t)p->Cout};
    cuuint32_t box[3] = {(cuuint32_t)kBlockK, (cuuint32_t)bn, 1};
    if ((rc = encode_map(&tBh, p->w_hi, brank, dims, strides, box, "B hi", p->plane_fmt))) return rc;
    if (split && (rc = encode_map(&tBl, p->w_lo, brank, dims, strides, box, "B lo", p->plane_fmt))) return rc;
  }
  if (!split) { tAl = tAh; tBl = tBh; }

  // split-K (see plan_splitk): only with a caller-provided workspace of the size shineon_conv2d_splitk_workspace_bytes says
  a.ksplit = 1;
  a.kb_per_split = a.num_kb;
  a.sk_ws = nullptr; a.sk_cnt = nullptr; a.sk_ctot = 0; a.sk_slice = 0;
  if (p->splitk_ws != nullptr && a.phases == 1) {
    const int n_tiles = cdiv(a.Cout, bn);
    int per = a.num_kb;
    const int ks = plan_splitk(m_tiles * n_tiles, a.num_kb, bn, a.chunk_kb, &per);
    if (ks > 1) {
      const size_t need = splitk_bytes(m_tiles * n_tiles, m_tiles, n_tiles, bn, ks);
      SHINEON_REQUIRE(p->splitk_ws_bytes >= need, "conv2d: split-K workspace of %zu bytes, need %zu", p->splitk_ws_bytes, need);
      SHINEON_REQUIRE((reinterpret_cast<uintptr_t>(p->splitk_ws) & 15) == 0, "conv2d: split-K workspace must be 16-byte aligned");
      a.ksplit = ks;
      a.kb_per_split = per;
      a.sk_cnt = reinterpret_cast<int*>(p->splitk_ws);
      a.sk_ws = reinterpret_cast<float*>(reinterpret_cast<char*>(p->splitk_ws) + (((size_t)m_tiles * n_tiles * sizeof(int) + 255) & ~(size_t)255));
      a.sk_ctot = n_tiles * bn;
      a.sk_slice = (long)m_tiles * kBlockM * a.sk_ctot;
    }
  }
  const bool multi = a.kb_per_split > a.chunk_kb;
  if (a.ksplit > 1) {  // split-K instantiations: slices never exceed one accumulation chunk (plan_splitk), BN 64 / 128 only
    if (bn == 64)
      return split ? launch_igemm<64, true, false, true>(tAh, tAl, tBh, tBl, a, m_tiles, p->stages, stream)
                   : launch_igemm<64, false, false, true>(tAh, tAl, tBh, tBl, a, m_tiles, p->stages, stream);
    return split ? launch_igemm<128, true, false, true>(tAh, tAl, tBh, tBl, a, m_tiles, p->stages, stream)
                 : launch_igemm<128, false, false, true>(tAh, tAl, tBh, tBl, a, m_tiles, p->stages, stream);
  }
#define SHINEON_LAUNCH(BN_)                                                                                                      \
  (split ? (multi ? launch_igemm<BN_, true, true>(tAh, tAl, tBh, tBl, a, m_tiles, p->stages, stream)                             \
                  : launch_igemm<BN_, true, false>(tAh, tAl, tBh, tBl, a, m_tiles, p->stages, stream))                           \
         : (multi ? launch_igemm<BN_, false, true>(tAh, tAl, tBh, tBl, a, m_tiles, p->stages, stream)                            \
                  : launch_igemm<BN_, false, false>(tAh, tAl, tBh, tBl, a, m_tiles, p->stages, stream)))
  switch (bn) {
    case 16: return SHINEON_LAUNCH(16);
    case 32: return SHINEON_LAUNCH(32);
    case 64: return SHINEON_LAUNCH(64);
    case 128: return SHINEON_LAUNCH(128);
    default: return SHINEON_LAUNCH(256);
  }
#undef SHINEON_LAUNCH
}

template <int BN, bool SPLIT>
static int launch_im2col(const CUtensorMap& tBh, const CUtensorMap& tBl, ConvArgs& a, const Im2colSrc& src, int m_tiles,
                         cudaStream_t stream) {
  constexpr int kBBytes = BN * kBlockK * 2;
  constexpr int kStageBytes = (SPLIT ? 2 : 1) * (kABytes + kBBytes);
  int stages = a.num_kb < 4 ? a.num_kb : 4;
  a.idesc = umma_idesc_f16(kBlockM, BN, a.fmt == SHINEON_FMT_FP16 ? 0 : 1);
  static int max_dyn_smem = -1;
  if (max_dyn_smem < 0) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, conv_im2col_igemm_kernel<BN, SPLIT>);
    if (e == cudaSuccess) {
      max_dyn_smem = 227 * 1024 - (int)fa.sharedSizeBytes;
      e = cudaFuncSetAttribute(conv_im2col_igemm_kernel<BN, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn_smem);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      max_dyn_smem = -1;
      return fail(SHINEON_ERR_CUDA, "conv_im2col shared-memory opt-in: %s", cudaGetErrorString(e));
    }
  }
  while (stages > 1 && stages * kStageBytes + 1024 > max_dyn_smem) --stages;
  a.stages = stages;
  a.n_tiles = cdiv(a.Cout, BN);
  a.total_tiles = m_tiles * a.n_tiles;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
  }
  const int grid = a.total_tiles < num_sms ? a.total_tiles : num_sms;
  klaunch(conv_im2col_igemm_kernel<BN, SPLIT>, grid, kI2cThreads, stages * kStageBytes + 1024, stream, tBh, tBl, a, src);
  return after_launch("conv_im2col_igemm_kernel");
}

extern "C" int shineon_conv2d_im2col_fwd(const shineon_conv2d_params* p, const float* x0, int C0, const float* x1, int C1,
                                         shineon_stream_t stream_) {
  SHINEON_REQUIRE(p && x0 && C0 > 0 && (x1 == nullptr) == (C1 == 0), "conv2d_im2col: bad input tensors");
  SHINEON_REQUIRE(p->w_hi && p->N > 0 && p->H > 0 && p->W > 0 && p->Cout > 0, "conv2d_im2col: bad shape");
  SHINEON_REQUIRE(p->cin_pad % kBlockK == 0 && p->cin_pad >= p->kh * p->kw * (C0 + C1), "conv2d_im2col: cin_pad (= padded K) %d", p->cin_pad);
  SHINEON_REQUIRE(p->Ho == (p->H + 2 * p->pad_h - p->kh) / p->stride + 1 && p->Wo == (p->W + 2 * p->pad_w - p->kw) / p->stride + 1, "conv2d_im2col: Ho/Wo");
  SHINEON_REQUIRE(p->y_f32 || p->y_hi, "conv2d_im2col: no output");
  SHINEON_REQUIRE((p->scale == nullptr) == (p->shift == nullptr), "conv2d_im2col: scale/shift");
  SHINEON_REQUIRE(p->plane_fmt == SHINEON_FMT_BF16 || p->plane_fmt == SHINEON_FMT_FP16, "conv2d_im2col: plane_fmt");
  ConvArgs a;
  a.N = p->N; a.Ho = p->Ho; a.Wo = p->Wo; a.Cout = p->Cout;
  a.kh = p->kh; a.kw = p->kw; a.stride = p->stride; a.pad_h = p->pad_h; a.pad_w = p->pad_w;
  a.cin_pad = p->cin_pad; a.cin_blocks = p->cin_pad / kBlockK; a.x_cstride = p->cin_pad;
  a.num_kb = a.cin_blocks;  // the GEMM is 1x1 over the padded K
  a.chunk_kb = a.num_kb;
  a.stages = 1; a.w_per_image = 0; a.stats = nullptr; a.phases = 1; a.phase_tiles = 0;
  a.ksplit = 1; a.kb_per_split = a.num_kb; a.sk_ws = nullptr; a.sk_cnt = nullptr; a.sk_ctot = 0; a.sk_slice = 0;
  a.bias = p->bias; a.scale = p->scale; a.shift = p->shift;
  a.pre_act = p->pre_act; a.post_act = p->post_act; a.act_param = p->act_param;
  a.fmt = p->plane_fmt; a.acc_scale = p->acc_scale == 0.f ? 1.f : p->acc_scale; a.idesc = 0;
  a.y_f32 = p->y_f32; a.y_hi = (plane_t*)p->y_hi; a.y_lo = (plane_t*)p->y_lo;
  a.oh_mul = 1; a.ow_mul = 1; a.oh_off = 0; a.ow_off = 0;
  a.out_H = p->Ho; a.out_W = p->Wo;
  a.out_cstride = p->out_cstride ? p->out_cstride : p->Cout;
  a.out_coffset = p->out_coffset;
  SHINEON_REQUIRE(a.out_coffset >= 0 && a.out_coffset + a.Cout <= a.out_cstride, "conv2d_im2col: output channel window out of range");
  pick_tile(a.N, a.Ho, a.Wo, a.nb, a.bh, a.bw);
  a.tiles_w = cdiv(a.Wo, a.bw);
  a.tiles_h = cdiv(a.Ho, a.bh);
  const int m_tiles = a.tiles_w * a.tiles_h * cdiv(a.N, a.nb);
  const bool split = p->w_lo != nullptr;
  int bn = p->tile_n;
  if (bn == 0) bn = a.Cout <= 16 ? 16 : a.Cout <= 32 ? 32 : a.Cout <= 64 ? 64 : 128;
  SHINEON_REQUIRE(bn == 16 || bn == 32 || bn == 64 || bn == 128, "conv2d_im2col: tile_n %d", bn);
  CUtensorMap tBh, tBl;
  int rc;
  {
    const cuuint64_t K = (cuuint64_t)p->cin_pad;
    cuuint64_t dims[2] = {K, (cuuint64_t)p->Cout};
    cuuint64_t strides[1] = {K * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)bn};
    if ((rc = encode_map(&tBh, p->w_hi, 2, dims, strides, box, "B hi", p->plane_fmt))) return rc;
    if (split && (rc = encode_map(&tBl, p->w_lo, 2, dims, strides, box, "B lo", p->plane_fmt))) return rc;
    if (!split) tBl = tBh;
  }
  Im2colSrc src{x0, x1, C0, C1, p->H, p->W};
  cudaStream_t stream = (cudaStream_t)stream_;
#define SHINEON_LAUNCH_I2C(BN_) (split ? launch_im2col<BN_, true>(tBh, tBl, a, src, m_tiles, stream) : launch_im2col<BN_, false>(tBh, tBl, a, src, m_tiles, stream))
  switch (bn) {
    case 16: return SHINEON_LAUNCH_I2C(16);
    case 32: return SHINEON_LAUNCH_I2C(32);
    case 64: return SHINEON_LAUNCH_I2C(64);
    default: return SHINEON_LAUNCH_I2C(128);
  }
#undef SHINEON_LAUNCH_I2C
}

extern "C" int shineon_conv2d_direct_fwd(const shineon_conv2d_params* p, shineon_stream_t stream) {
  ConvArgs a;
  int rc = fill_args(p, a);
  if (rc) return rc;
  SHINEON_REQUIRE(a.phases == 1, "conv2d_direct: deconv_phases is a tcgen05-kernel mode (run the four phases separately)");
  long total = (long)a.N * a.Ho * a.Wo * a.Cout;
  long blocks = (total + 127) / 128;
  if (blocks > 148 * 64) blocks = 148 * 64;
  klaunch(conv_direct_kernel, (int)blocks, 128, 0, (cudaStream_t)stream, 
      (const plane_t*)p->x_hi, (const plane_t*)p->x_lo, (const plane_t*)p->w_hi,
      (const plane_t*)p->w_lo, p->H, p->W, a);
  return after_launch("conv_direct_kernel");
}

extern "C" int shineon_pack_conv_weight(const float* w, void* w_hi, void* w_lo, int Cout, int Cin, int kh, int kw,
                                        int cin_pad, const int32_t* chan_map, int transpose_io, int plane_fmt,
                                        float w_scale, shineon_stream_t stream) {
  SHINEON_REQUIRE(plane_fmt == SHINEON_FMT_BF16 || plane_fmt == SHINEON_FMT_FP16, "pack_conv_weight: plane_fmt %d", plane_fmt);
  if (w_scale == 0.f) w_scale = 1.f;
  SHINEON_REQUIRE(w && w_hi, "pack_conv_weight: null pointer");
  SHINEON_REQUIRE(Cout > 0 && Cin > 0 && kh > 0 && kw > 0 && cin_pad >= 1, "pack_conv_weight: bad shape");
  SHINEON_REQUIRE(chan_map != nullptr || cin_pad >= Cin, "pack_conv_weight: cin_pad < Cin");
  const dim3 grid(cdiv(cin_pad, 128), Cout < 65535 ? Cout : 65535);
  klaunch(pack_conv_weight_kernel, grid, 128, 0, (cudaStream_t)stream, w, (plane_t*)w_hi, (plane_t*)w_lo, Cout, Cin, kh, kw, cin_pad,
                                                                  chan_map, transpose_io, plane_fmt, w_scale);
  return after_launch("pack_conv_weight_kernel");
}

extern "C" int shineon_pack_deconv4x4s2_weight(const float* w, void* w_hi, void* w_lo, int Cin, int Cout, int cin_pad,
                                               const int32_t* chan_map, int plane_fmt, float w_scale, shineon_stream_t stream) {
  SHINEON_REQUIRE(plane_fmt == SHINEON_FMT_BF16 || plane_fmt == SHINEON_FMT_FP16, "pack_deconv4x4s2_weight: plane_fmt %d", plane_fmt);
  SHINEON_REQUIRE(w && w_hi && Cin > 0 && Cout > 0 && cin_pad >= 1, "pack_deconv4x4s2_weight: bad arguments");
  SHINEON_REQUIRE(chan_map != nullptr || cin_pad >= Cin, "pack_deconv4x4s2_weight: cin_pad < Cin");
  if (w_scale == 0.f) w_scale = 1.f;
  long total = 16l * Cout * cin_pad;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  klaunch(pack_deconv4x4s2_weight_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, w, (plane_t*)w_hi, (plane_t*)w_lo, Cout, Cin,
                                                                              cin_pad, chan_map, plane_fmt, w_scale);
  return after_launch("pack_deconv4x4s2_weight_kernel");
}
