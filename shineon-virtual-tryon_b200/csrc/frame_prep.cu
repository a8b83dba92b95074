// Dataset-side per-frame tensor prep on the GPU (reference: datasets/tryon_dataset.py).
// The reference builds every network input on the CPU inside Dataset.__getitem__ from decoded 8-bit images:
//   ToTensor + Normalize(0.5, 0.5)                       tryon_dataset.py:109-119,149-152
//   cloth mask   where(cloth >= thr, 0, 1)[0]            :168-175   (compares the NORMALISED cloth, as written)
//   head         im * phead - (1 - phead)                :323-344   (phead = LIP label in a fixed set)
//   silhouette   (parse > 0) * 255 -> PIL BILINEAR /16 -> PIL BILINEAR x16 -> normalise          :346-367
//   cocopose     18 channels, constant -1 as written (:415-423) + the square visualisation        :389-447
//   .flo         Middlebury flow bytes -> (x - 0.5) / 0.5                    flow_utils.py:7-26, tryon_dataset.py:288-289
// Moving this to the device lets a frame cross PCIe as 10 bytes/pixel of uint8 instead of 112 bytes/pixel of f32.
// Byte / integer work, bit-exact against the reference: the PIL resize is done with Pillow's own fixed-point scheme
// (Resample.c: 22-bit coefficients, rounding to uint8 after each pass), coefficients from shineon_pil_bilinear_coeffs.
#include <cmath>
#include <type_traits>

#include "common.cuh"

namespace shineon {

__device__ __forceinline__ float norm_u8(uint8_t u) {
  // ToTensor: float(u) / 255 (IEEE division), Normalize: (t - 0.5) / 0.5
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.f), 0.5f), 0.5f);
}

// LIP labels kept by get_person_head: HAT 1, HAIR 2, SUNGLASSES 4, SOCKS 8, PANTS 9, SCARF 11, SKIRT 12, FACE 13,
// LEFT_LEG 16, RIGHT_LEG 17, LEFT_SHOE 18, RIGHT_SHOE 19
constexpr uint32_t kHeadMask = (1u << 1) | (1u << 2) | (1u << 4) | (1u << 8) | (1u << 9) | (1u << 11) | (1u << 12) |
                               (1u << 13) | (1u << 16) | (1u << 17) | (1u << 18) | (1u << 19);

struct PrepArgs {
  const uint8_t* image;      // [F,H,W,3] or null
  const uint8_t* parse;      // [F,H,W] or null (needed for head)
  const uint8_t* cloth;      // [F,H,W,3] or null
  const uint8_t* densepose;  // [F,H,W,3] or null
  float* image_out;          // [F,3,H,W] or null
  float* cloth_out;          // [F,3,H,W]
  float* cloth_mask_out;     // [F,1,H,W]
  float* densepose_out;      // [F,3,H,W]
  float* agnostic_out;       // [F,4,H,W]: channel 0 = silhouette (other kernel), 1..3 = head
  float* cocopose_out;       // [F,J,H,W] constant -1
  int HW, J;
  float cloth_thr;
};

// 4 consecutive pixels per thread: 12-byte (3 x uchar4) reads of the interleaved images, float4 stores per channel plane
__global__ void __launch_bounds__(256) frame_prep_pointwise_kernel(const PrepArgs a) {
  pdl_grid_sync();
  const int f = blockIdx.y;
  const int p0 = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (p0 >= a.HW) return;
  const long fo = (long)f * a.HW;
  auto load12 = [&](const uint8_t* base, uint8_t (&v)[12]) {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(base + (fo + p0) * 3);  // (fo + p0) * 3 is a multiple of 4
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const uint32_t w = __ldg(s + i);
      v[4 * i] = w & 0xff; v[4 * i + 1] = (w >> 8) & 0xff; v[4 * i + 2] = (w >> 16) & 0xff; v[4 * i + 3] = w >> 24;
    }
  };
  auto store_rgb = [&](float* out, const float (&v)[12]) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      *reinterpret_cast<float4*>(out + ((long)f * 3 + c) * a.HW + p0) = make_float4(v[c], v[3 + c], v[6 + c], v[9 + c]);
  };
  uint8_t u[12];
  float v[12];
  if (a.cloth) {
    load12(a.cloth, u);
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i] = norm_u8(u[i]);
    store_rgb(a.cloth_out, v);
    if (a.cloth_mask_out)  // channel 0 only (tryon_dataset.py:174)
      *reinterpret_cast<float4*>(a.cloth_mask_out + fo + p0) =
          make_float4(v[0] >= a.cloth_thr ? 0.f : 1.f, v[3] >= a.cloth_thr ? 0.f : 1.f, v[6] >= a.cloth_thr ? 0.f : 1.f,
                      v[9] >= a.cloth_thr ? 0.f : 1.f);
  }
  if (a.densepose) {
    load12(a.densepose, u);
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i] = norm_u8(u[i]);
    store_rgb(a.densepose_out, v);
  }
  if (a.image) {
    load12(a.image, u);
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i] = norm_u8(u[i]);
    if (a.image_out) store_rgb(a.image_out, v);
    if (a.agnostic_out) {
      const uint32_t pw = __ldg(reinterpret_cast<const uint32_t*>(a.parse + fo + p0));
      float ph[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t lab = (pw >> (8 * i)) & 0xff;
        ph[i] = (lab < 32 && ((kHeadMask >> lab) & 1u)) ? 1.f : 0.f;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __fsub_rn(__fmul_rn(v[3 * i + c], ph[i]), __fsub_rn(1.f, ph[i]));  // im*phead - (1-phead)
        *reinterpret_cast<float4*>(a.agnostic_out + ((long)f * 4 + 1 + c) * a.HW + p0) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  if (a.cocopose_out) {
    const float4 m1 = make_float4(-1.f, -1.f, -1.f, -1.f);
    for (int j = 0; j < a.J; ++j) *reinterpret_cast<float4*>(a.cocopose_out + ((long)f * a.J + j) * a.HW + p0) = m1;
  }
}

// Pillow's 8-bit resampling pass: out = clip8(((1 << 21) + sum px * k) >> 22)
__device__ __forceinline__ uint8_t pil_clip8(int ss) {
  ss >>= 22;
  return (uint8_t)(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
}

struct ResizeTab {
  const int* bounds;  // [out][2] = (first input index, tap count)
  const int* kk;      // [out][ksize]
  int ksize;
};

// One CTA per frame; the three intermediate images live in shared memory (H*w16 + h16*w16 + h16*W bytes).
__global__ void __launch_bounds__(1024)
    body_silhouette_kernel(const uint8_t* __restrict__ parse, float* __restrict__ out, long out_frame_stride,
                           uint8_t* __restrict__ out_u8, int H, int W, ResizeTab dw, ResizeTab dh, ResizeTab uw, ResizeTab uh) {
  pdl_grid_sync();
  extern __shared__ uint8_t sm[];
  const int w16 = W / 16, h16 = H / 16;
  uint8_t* tA = sm;                 // [H][w16]   after the horizontal down pass
  uint8_t* tS = tA + H * w16;       // [h16][w16] after the vertical down pass
  uint8_t* tB = tS + h16 * w16;     // [h16][W]   after the horizontal up pass
  const uint8_t* src = parse + (long)blockIdx.x * H * W;
  const int nt = blockDim.x;  // 1024: one CTA per frame, a single wave -- the kernel's time is one CTA's latency
  for (int e = threadIdx.x; e < H * w16; e += nt) {
    const int y = e / w16, xx = e - y * w16;
    const int x0 = dw.bounds[2 * xx], cnt = dw.bounds[2 * xx + 1];
    int ss = 1 << 21;
    for (int t = 0; t < cnt; ++t) ss += (src[y * W + x0 + t] > 0 ? 255 : 0) * dw.kk[xx * dw.ksize + t];
    tA[e] = pil_clip8(ss);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < h16 * w16; e += nt) {
    const int yy = e / w16, xx = e - yy * w16;
    const int y0 = dh.bounds[2 * yy], cnt = dh.bounds[2 * yy + 1];
    int ss = 1 << 21;
    for (int t = 0; t < cnt; ++t) ss += tA[(y0 + t) * w16 + xx] * dh.kk[yy * dh.ksize + t];
    tS[e] = pil_clip8(ss);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < h16 * W; e += nt) {
    const int yy = e / W, x = e - yy * W;
    const int x0 = uw.bounds[2 * x], cnt = uw.bounds[2 * x + 1];
    int ss = 1 << 21;
    for (int t = 0; t < cnt; ++t) ss += tS[yy * w16 + x0 + t] * uw.kk[x * uw.ksize + t];
    tB[e] = pil_clip8(ss);
  }
  __syncthreads();
  float* dst = out ? out + (long)blockIdx.x * out_frame_stride : nullptr;
  uint8_t* dst8 = out_u8 ? out_u8 + (long)blockIdx.x * H * W : nullptr;
  for (int e = threadIdx.x; e < H * W; e += nt) {
    const int y = e / W, x = e - y * W;
    const int y0 = uh.bounds[2 * y], cnt = uh.bounds[2 * y + 1];
    int ss = 1 << 21;
    for (int t = 0; t < cnt; ++t) ss += tB[(y0 + t) * W + x] * uh.kk[y * uh.ksize + t];
    const uint8_t r = pil_clip8(ss);
    if (dst) dst[e] = norm_u8(r);
    if (dst8) dst8[e] = r;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Frame prep written straight into the first layers' operands (TryOnPipeline.run_raw).  Per frame it produces, from the
// decoded 8-bit images and the 8-bit silhouette of body_silhouette_kernel,
//   pg  [F, H/2+1, W/2+1, cg]: shifted space-to-depth planes of cat(agnostic, cocopose)  (WarpModel's person input),
//   pu  [F, H/2+1, W/2+1, cu]: the same for cat(agnostic, densepose, cloth'): the U-Net stem's input; the three cloth'
//                              channels are left zero here and filled by tps_warp_u8_planes_kernel,
//   pc  [F, H/2,   W/2,   cc]: im2col planes (4x4, stride 2, pad 1; k = (fy*4+fx)*3 + c) of the cloth (extractionB's stem),
// each as 16-bit hi (+ lo) halves -- bit-identical to shineon_frame_prep -> torch.cat -> shineon_nchw_s2d_planes /
// shineon_nchw_im2col_planes, which this replaces: 1.0 GB written per 80-frame step instead of 2.9 GB moved.
// One CTA = 32 consecutive X of one (frame, Y): values are assembled as finished plane rows in shared memory, then
// copied out with fully coalesced 16-byte stores.
struct PrepPlanesArgs {
  const uint8_t *image, *parse, *cloth, *densepose, *sil;
  plane_t *pg_hi, *pg_lo, *pu_hi, *pu_lo, *pc_hi, *pc_lo;
  int H, W, J, fmt;
};
constexpr int kPrepPx = 32;

// CG / CU / CC: channel pads of the three outputs (compile-time: index math without divisions).  ToTensor + Normalize of
// a byte has 256 possible results: a shared table (hi/lo halves precomputed) replaces two IEEE divisions and a split per use.
template <int CG, int CU, int CC>
__global__ void __launch_bounds__(256) frame_prep_planes_kernel(const PrepPlanesArgs a) {
  pdl_grid_sync();
  constexpr int ROW = CG + CU + CC;  // halfwords per pixel and half (hi or lo)
  extern __shared__ __align__(16) plane_t sm_rows[];
  __shared__ uint32_t s_lut[256];   // hi | lo << 16 of norm_u8(b)
  __shared__ float s_lutf[256];
  const int f = blockIdx.z, Y = blockIdx.y, X0 = blockIdx.x * kPrepPx;
  const int Hz = a.H / 2 + 1, Wz = a.W / 2 + 1;
  plane_t* s_hi = sm_rows;                     // [kPrepPx][CG | CU | CC]
  plane_t* s_lo = sm_rows + kPrepPx * ROW;
  {
    const float v = norm_u8((uint8_t)threadIdx.x);
    plane_t h, l;
    split16(v, a.fmt, h, l);
    s_lut[threadIdx.x] = (uint32_t)h | ((uint32_t)l << 16);
    s_lutf[threadIdx.x] = v;
    uint4* z = reinterpret_cast<uint4*>(sm_rows);  // zero: padding channels, out-of-image taps, the cloth' slots
#pragma unroll
    for (int i = 0; i < 2 * kPrepPx * ROW / 8 / 256; ++i) z[threadIdx.x + i * 256] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  const long fo = (long)f * a.H * a.W;
  const int cper_g = 4 + a.J;
  constexpr int cper_u = 10;
  plane_t neg1_hi, neg1_lo;
  split16(-1.f, a.fmt, neg1_hi, neg1_lo);  // -1 is exact in both formats: lo == 0
  if (threadIdx.x < 4 * kPrepPx) {  // (pixel, position q): the person channels of pg and pu
    const int px = threadIdx.x >> 2, q = threadIdx.x & 3;
    const int X = X0 + px;
    const int iy = 2 * Y - 1 + (q >> 1), ix = 2 * X - 1 + (q & 1);
    if (X < Wz && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) {
      const long p = fo + (long)iy * a.W + ix;
      const uint32_t lab = a.parse[p];
      const bool head = lab < 32 && ((kHeadMask >> lab) & 1u);
      plane_t* gh = s_hi + px * ROW + q * cper_g;
      plane_t* gl = s_lo + px * ROW + q * cper_g;
      plane_t* uh = s_hi + px * ROW + CG + q * cper_u;
      plane_t* ul = s_lo + px * ROW + CG + q * cper_u;
      const uint32_t sv = s_lut[a.sil[p]];
      gh[0] = uh[0] = (plane_t)sv;
      gl[0] = ul[0] = (plane_t)(sv >> 16);
#pragma unroll
      for (int c = 0; c < 3; ++c) {  // im * phead - (1 - phead): the pixel itself inside the head labels, -1 elsewhere
        const uint32_t hv = head ? s_lut[a.image[p * 3 + c]] : (uint32_t)neg1_hi;
        gh[1 + c] = uh[1 + c] = (plane_t)hv;
        gl[1 + c] = ul[1 + c] = (plane_t)(hv >> 16);
        const uint32_t dv = s_lut[a.densepose[p * 3 + c]];
        uh[4 + c] = (plane_t)dv;
        ul[4 + c] = (plane_t)(dv >> 16);
      }
      for (int j = 0; j < a.J; ++j) gh[4 + j] = neg1_hi;  // cocopose maps: constant -1 as the reference writes them (lo = 0)
    }
  }
  for (int it = threadIdx.x; it < 16 * kPrepPx; it += 256) {  // (pixel, filter tap): the cloth's im2col row
    const int px = it >> 4, tap = it & 15;
    const int X = X0 + px;
    const int iy = 2 * Y - 1 + (tap >> 2), ix = 2 * X - 1 + (tap & 3);
    if (Y < a.H / 2 && X < a.W / 2 && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) {
      const long p = fo + (long)iy * a.W + ix;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const uint32_t cv = s_lut[a.cloth[p * 3 + c]];
        s_hi[px * ROW + CG + CU + tap * 3 + c] = (plane_t)cv;
        s_lo[px * ROW + CG + CU + tap * 3 + c] = (plane_t)(cv >> 16);
      }
    }
  }
  __syncthreads();
  // coalesced copy-out: per destination tensor, 16-byte units of the CTA's contiguous row segment
  const int nz = min(kPrepPx, Wz - X0), nc = (Y < a.H / 2) ? max(0, min(kPrepPx, a.W / 2 - X0)) : 0;
  auto copy_out = [&](plane_t* dst_hi, plane_t* dst_lo, auto cw_tag, int soff, long row0, int npx) {
    constexpr int CW = decltype(cw_tag)::value;
    constexpr int UPP = CW / 8;  // 16-byte units per pixel
    if (!dst_hi || npx <= 0) return;
    for (int i = threadIdx.x; i < npx * UPP; i += 256) {
      const int px = i / UPP, u = i % UPP;
      *reinterpret_cast<uint4*>(dst_hi + (row0 + px) * CW + u * 8) = *reinterpret_cast<const uint4*>(s_hi + px * ROW + soff + u * 8);
      if (dst_lo)
        *reinterpret_cast<uint4*>(dst_lo + (row0 + px) * CW + u * 8) = *reinterpret_cast<const uint4*>(s_lo + px * ROW + soff + u * 8);
    }
  };
  const long zrow = ((long)f * Hz + Y) * Wz + X0;
  copy_out(a.pg_hi, a.pg_lo, std::integral_constant<int, CG>{}, 0, zrow, nz);
  copy_out(a.pu_hi, a.pu_lo, std::integral_constant<int, CU>{}, CG, zrow, nz);
  copy_out(a.pc_hi, a.pc_lo, std::integral_constant<int, CC>{}, CG + CU, ((long)f * (a.H / 2) + Y) * (a.W / 2) + X0, nc);
}

// im_cocopose: union of the filled squares ImageDraw.rectangle((x-r, y-r, x+r, y+r)) draws for joints with x > 1, y > 1
__global__ void __launch_bounds__(256)
    cocopose_vis_kernel(const double* __restrict__ pose, float* __restrict__ out, int H, int W, int J, int radius) {
  pdl_grid_sync();
  __shared__ int box[64][4];
  const int f = blockIdx.y;
  for (int j = threadIdx.x; j < J; j += 256) {
    const double x = pose[((long)f * J + j) * 3], y = pose[((long)f * J + j) * 3 + 1];
    const bool on = x > 1.0 && y > 1.0;
    box[j][0] = on ? (int)(x - radius) : 1;  // C truncation of the double, like the drawing code
    box[j][1] = on ? (int)(y - radius) : 1;
    box[j][2] = on ? (int)(x + radius) : 0;  // empty box when the joint is absent
    box[j][3] = on ? (int)(y + radius) : 0;
  }
  __syncthreads();
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= H * W) return;
  const int y = p / W, x = p - y * W;
  bool hit = false;
  for (int j = 0; j < J; ++j) hit |= (x >= box[j][0] && x <= box[j][2] && y >= box[j][1] && y <= box[j][3]);
  out[(long)f * H * W + p] = hit ? 1.f : -1.f;  // norm_u8(255) = 1, norm_u8(0) = -1
}

__global__ void __launch_bounds__(256)
    flo_decode_kernel(const float* __restrict__ uv, float* __restrict__ out, long hw) {
  pdl_grid_sync();
  const long p = (long)blockIdx.x * 256 + threadIdx.x;
  if (p >= hw) return;
  const float u = __ldg(uv + 2 * p), v = __ldg(uv + 2 * p + 1);  // the payload starts 12 bytes into the file: 4-byte aligned only
  out[p] = __fdiv_rn(__fsub_rn(u, 0.5f), 0.5f);
  out[hw + p] = __fdiv_rn(__fsub_rn(v, 0.5f), 0.5f);
}

}  // namespace shineon

using namespace shineon;

// Host-side: Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR filter (support 1.0).
// bounds: 2*out_size ints; kk: out_size*ksize ints with ksize = the return value; pass kk = NULL to query ksize.
extern "C" int shineon_pil_bilinear_coeffs(int in_size, int out_size, int* bounds, int* kk) {
  if (in_size <= 0 || out_size <= 0) return fail(SHINEON_ERR_ARG, "pil_bilinear_coeffs: bad sizes %d -> %d", in_size, out_size);
  double scale = (double)((float)in_size - 0.0f) / out_size;
  double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  if (!kk || !bounds) return ksize;
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double w[1024];
    if (xmax > 1024) return fail(SHINEON_ERR_UNSUPPORTED, "pil_bilinear_coeffs: filter too wide");
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double t = (x + xmin - center + 0.5) * ss;
      if (t < 0.0) t = -t;
      w[x] = t < 1.0 ? 1.0 - t : 0.0;
      ww += w[x];
    }
    for (int x = 0; x < ksize; ++x) {
      int q = 0;
      if (x < xmax) {
        double v = w[x];
        if (ww != 0.0) v /= ww;
        q = v < 0 ? (int)(-0.5 + v * (1 << 22)) : (int)(0.5 + v * (1 << 22));
      }
      kk[xx * ksize + x] = q;
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  return ksize;
}

extern "C" int shineon_frame_prep(const shineon_frame_prep_params* p, shineon_stream_t stream) {
  SHINEON_REQUIRE(p != nullptr, "frame_prep: null params");
  SHINEON_REQUIRE(p->F > 0 && p->F <= 65535 && p->H > 0 && p->W > 0, "frame_prep: bad shape");
  SHINEON_REQUIRE((p->H * p->W) % 4 == 0, "frame_prep: H*W must be a multiple of 4");
  SHINEON_REQUIRE(!p->cloth || p->cloth_out, "frame_prep: cloth given without cloth_out");
  SHINEON_REQUIRE(!p->densepose || p->densepose_out, "frame_prep: densepose given without densepose_out");
  SHINEON_REQUIRE(!p->agnostic_out || (p->image && p->parse), "frame_prep: agnostic needs image and parse");
  SHINEON_REQUIRE(!p->cocopose_out || (p->n_joints > 0 && p->n_joints <= 64), "frame_prep: n_joints must be in 1..64");
  SHINEON_REQUIRE(!p->im_cocopose_out || (p->pose && p->n_joints > 0 && p->n_joints <= 64), "frame_prep: im_cocopose needs pose");
  SHINEON_REQUIRE(((reinterpret_cast<uintptr_t>(p->image) | reinterpret_cast<uintptr_t>(p->parse) | reinterpret_cast<uintptr_t>(p->cloth) |
                    reinterpret_cast<uintptr_t>(p->densepose)) & 3) == 0, "frame_prep: uint8 inputs must be 4-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = p->H * p->W;
  PrepArgs a;
  a.image = p->image; a.parse = p->parse; a.cloth = p->cloth; a.densepose = p->densepose;
  a.image_out = p->image_out; a.cloth_out = p->cloth_out; a.cloth_mask_out = p->cloth_mask_out;
  a.densepose_out = p->densepose_out; a.agnostic_out = p->agnostic_out; a.cocopose_out = p->cocopose_out;
  a.HW = HW; a.J = p->n_joints; a.cloth_thr = p->cloth_mask_threshold;
  klaunch(frame_prep_pointwise_kernel, dim3(cdiv(HW / 4, 256), p->F), 256, 0, st, a);
  int rc = after_launch("frame_prep_pointwise_kernel");
  if (rc) return rc;
  if (p->agnostic_out) {
    SHINEON_REQUIRE(p->H % 16 == 0 && p->W % 16 == 0, "frame_prep: silhouette needs H, W multiples of 16");
    SHINEON_REQUIRE(p->tab_bounds[0] && p->tab_kk[0] && p->tab_bounds[1] && p->tab_kk[1] && p->tab_bounds[2] && p->tab_kk[2] &&
                        p->tab_bounds[3] && p->tab_kk[3], "frame_prep: resize tables missing (shineon_pil_bilinear_coeffs)");
    ResizeTab t[4];
    for (int i = 0; i < 4; ++i) t[i] = ResizeTab{p->tab_bounds[i], p->tab_kk[i], p->tab_ksize[i]};
    const size_t smem = (size_t)p->H * (p->W / 16) + (size_t)(p->H / 16) * (p->W / 16) + (size_t)(p->H / 16) * p->W;
    SHINEON_REQUIRE(smem <= 48 * 1024, "frame_prep: frame too large for the silhouette kernel");
    klaunch(body_silhouette_kernel, p->F, 1024, smem, st, p->parse, p->agnostic_out, (long)4 * HW, nullptr, p->H, p->W, t[0], t[1], t[2], t[3]);
    rc = after_launch("body_silhouette_kernel");
    if (rc) return rc;
  }
  if (p->im_cocopose_out) {
    klaunch(cocopose_vis_kernel, dim3(cdiv(HW, 256), p->F), 256, 0, st, p->pose, p->im_cocopose_out, p->H, p->W, p->n_joints, p->radius);
    rc = after_launch("cocopose_vis_kernel");
    if (rc) return rc;
  }
  return SHINEON_OK;
}

extern "C" int shineon_frame_prep_planes(const shineon_frame_prep_planes_params* p, shineon_stream_t stream) {
  SHINEON_REQUIRE(p != nullptr, "frame_prep_planes: null params");
  SHINEON_REQUIRE(p->F > 0 && p->F <= 65535 && p->H > 0 && p->W > 0 && p->H % 16 == 0 && p->W % 16 == 0,
                  "frame_prep_planes: bad shape (H, W multiples of 16)");
  SHINEON_REQUIRE(p->image && p->parse && p->cloth && p->densepose && p->silhouette_scratch, "frame_prep_planes: null input");
  SHINEON_REQUIRE(p->gmm_hi && p->unet_hi && p->cloth_hi, "frame_prep_planes: null output");
  SHINEON_REQUIRE((p->gmm_lo == nullptr) == (p->unet_lo == nullptr) && (p->gmm_lo == nullptr) == (p->cloth_lo == nullptr),
                  "frame_prep_planes: lo planes must be all present or all absent");
  SHINEON_REQUIRE(p->plane_fmt == SHINEON_FMT_BF16 || p->plane_fmt == SHINEON_FMT_FP16, "frame_prep_planes: plane_fmt %d", p->plane_fmt);
  SHINEON_REQUIRE(p->n_joints >= 0 && p->n_joints <= 64, "frame_prep_planes: n_joints");
  SHINEON_REQUIRE(p->gmm_cpad % 8 == 0 && p->gmm_cpad >= 4 * (4 + p->n_joints) && p->unet_cpad % 8 == 0 && p->unet_cpad >= 40 &&
                      p->cloth_cpad % 8 == 0 && p->cloth_cpad >= 48, "frame_prep_planes: channel pads too small / not multiples of 8");
  SHINEON_REQUIRE(p->tab_bounds[0] && p->tab_kk[0] && p->tab_bounds[1] && p->tab_kk[1] && p->tab_bounds[2] && p->tab_kk[2] &&
                      p->tab_bounds[3] && p->tab_kk[3], "frame_prep_planes: resize tables missing (shineon_pil_bilinear_coeffs)");
  cudaStream_t st = (cudaStream_t)stream;
  ResizeTab t[4];
  for (int i = 0; i < 4; ++i) t[i] = ResizeTab{p->tab_bounds[i], p->tab_kk[i], p->tab_ksize[i]};
  const size_t smem_s = (size_t)p->H * (p->W / 16) + (size_t)(p->H / 16) * (p->W / 16) + (size_t)(p->H / 16) * p->W;
  SHINEON_REQUIRE(smem_s <= 48 * 1024, "frame_prep_planes: frame too large for the silhouette kernel");
  klaunch(body_silhouette_kernel, p->F, 1024, smem_s, st, p->parse, nullptr, 0, p->silhouette_scratch, p->H, p->W, t[0], t[1], t[2], t[3]);
  int rc = after_launch("body_silhouette_kernel");
  if (rc) return rc;
  PrepPlanesArgs a;
  a.image = p->image; a.parse = p->parse; a.cloth = p->cloth; a.densepose = p->densepose; a.sil = p->silhouette_scratch;
  a.pg_hi = (plane_t*)p->gmm_hi; a.pg_lo = (plane_t*)p->gmm_lo; a.pu_hi = (plane_t*)p->unet_hi; a.pu_lo = (plane_t*)p->unet_lo;
  a.pc_hi = (plane_t*)p->cloth_hi; a.pc_lo = (plane_t*)p->cloth_lo;
  a.H = p->H; a.W = p->W; a.J = p->n_joints; a.fmt = p->plane_fmt;
  dim3 grid(cdiv(p->W / 2 + 1, kPrepPx), p->H / 2 + 1, p->F);
  // channel pads are compile-time: the recipe's (18 joints: 88 -> 128, 40 -> 64, 48 -> 64) and the joint-free variant
  if (p->gmm_cpad == 128 && p->unet_cpad == 64 && p->cloth_cpad == 64) {
    constexpr size_t smem = (size_t)2 * kPrepPx * (128 + 64 + 64) * sizeof(plane_t);
    klaunch(frame_prep_planes_kernel<128, 64, 64>, grid, 256, smem, st, a);
  } else if (p->gmm_cpad == 64 && p->unet_cpad == 64 && p->cloth_cpad == 64) {
    constexpr size_t smem = (size_t)2 * kPrepPx * (64 + 64 + 64) * sizeof(plane_t);
    klaunch(frame_prep_planes_kernel<64, 64, 64>, grid, 256, smem, st, a);
  } else {
    return fail(SHINEON_ERR_UNSUPPORTED, "frame_prep_planes: channel pads (%d, %d, %d) not compiled (128/64/64 and 64/64/64 are)",
                p->gmm_cpad, p->unet_cpad, p->cloth_cpad);
  }
  return after_launch("frame_prep_planes_kernel");
}

extern "C" int shineon_flo_decode(const void* flo_payload, float* out, int H, int W, shineon_stream_t stream) {
  SHINEON_REQUIRE(flo_payload && out && H > 0 && W > 0, "flo_decode: bad arguments");
  SHINEON_REQUIRE((reinterpret_cast<uintptr_t>(flo_payload) & 3) == 0, "flo_decode: payload must be 4-byte aligned");
  const long hw = (long)H * W;
  klaunch(flo_decode_kernel, (unsigned)((hw + 255) / 256), 256, 0, (cudaStream_t)stream, (const float*)flo_payload, out, hw);
  return after_launch("flo_decode_kernel");
}
