/*
 * shineon_b200.h — C ABI of libshineon_b200.so (sm_100a).
 *
 * Drop-in boundary for the per-frame try-on hot path of
 * andrewjong/ShineOn-Virtual-Tryon.  Every entry point is `extern "C"`, takes
 * plain device pointers + sizes + a cudaStream_t, never allocates, never
 * synchronises, only enqueues on the given stream, and returns 0 on success or
 * a negative shineon_status (the reference's extensions return `int 1` and
 * swallow launch errors: correlation_cuda.cc:80-87, resample2d_cuda.cc:6-13;
 * this ABI is stricter on purpose).  `shineon_last_error()` returns the text
 * of the last failure on the calling thread.
 *
 * Reference interfaces each group replaces (paths relative to the reference
 * checkout):
 *   resample2d   models/flownet2_pytorch/networks/resample2d_package/resample2d_cuda.cc:6-31
 *   channelnorm  models/flownet2_pytorch/networks/channelnorm_package/channelnorm_cuda.cc:6-30
 *   correlation  models/flownet2_pytorch/networks/correlation_package/correlation_cuda.cc:10-172
 *   tps / grid_sample   models/networks/cpvton/warp.py:116-318, models/warp_model.py:85-86,143-145
 *   conv2d_igemm / pack / instnorm / prep / attention / compose
 *                models/networks/cpvton/{warp,unet}.py, models/networks/attention/sagan.py,
 *                models/unet_mask_model.py:64-135 (stock torch.nn ops there)
 *
 * Layout conventions
 *   "NCHW f32"   : the reference's public tensor layout (contiguous).
 *   "NHWC planes": internal activation layout of the tensor-core path:
 *                  two 16-bit tensors [N,H,W,Cpad] (hi, lo; bf16 or fp16, see shineon_plane_fmt) with x ~= hi + lo
 *                  (lo may be NULL in single-bf16 mode); Cpad % 64 == 0,
 *                  channels >= C are zero.
 */
#ifndef SHINEON_B200_H_
#define SHINEON_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* shineon_stream_t; /* cudaStream_t */

enum shineon_status {
  SHINEON_OK = 0,
  SHINEON_ERR_ARG = -1,         /* bad argument (shape, alignment, null pointer) */
  SHINEON_ERR_CUDA = -2,        /* a CUDA runtime/driver call failed */
  SHINEON_ERR_UNSUPPORTED = -3, /* valid request this build has no kernel for */
};

enum shineon_act {
  SHINEON_ACT_NONE = 0,
  SHINEON_ACT_RELU = 1,
  SHINEON_ACT_LEAKY = 2, /* slope in act_param */
  SHINEON_ACT_GELU = 3,  /* exact erf form (nn.GELU()) */
  SHINEON_ACT_SWISH = 4, /* x*sigmoid(x)  activation.py:13-18 */
  SHINEON_ACT_SINE = 5,  /* sin(30x)      activation.py:4-11 */
  SHINEON_ACT_TANH = 6,
  SHINEON_ACT_SIGMOID = 7,
};

enum shineon_padding_mode { SHINEON_PAD_ZEROS = 0, SHINEON_PAD_BORDER = 1 };

/* Encoding of the 16-bit hi/lo activation / weight planes (argument `plane_fmt`):
 * BF16: range-safe, hi+lo = 16 mantissa bits;  FP16: hi+lo = 22 mantissa bits (fp32-grade products), values
 * are clamped to +-65000 when written (weights are pre-scaled by a power of two, see acc_scale). */
enum shineon_plane_fmt { SHINEON_FMT_BF16 = 0, SHINEON_FMT_FP16 = 1 };

int shineon_version(void);
const char* shineon_last_error(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
uint64_t shineon_launch_count(void);

/* ------------------------------------------------------------------ */
/* G5/G6: thin-plate-spline grid + bilinear grid_sample                */
/* ------------------------------------------------------------------ */

/* Constant tables of TpsGridGen (all f32, device memory), built by the host exactly like the
 * reference constructor does (warp.py:116-157): Li [(gs*gs+3)^2] = compute_L_inverse (warp.py:169-189),
 * P_X/P_Y [gs*gs] control points, grid_X [W] / grid_Y [H] = float32(np.linspace(-1,1,.)). */
typedef struct shineon_tps_tables {
  const float* Li;
  const float* P_X;
  const float* P_Y;
  const float* grid_X;
  const float* grid_Y;
  int grid_size;
} shineon_tps_tables;

/* TpsGridGen.forward (warp.py:159-318).  theta [B, 2*gs*gs] f32 -> grid [B,H,W,2] f32. */
int shineon_tps_grid_fwd(const float* theta, const shineon_tps_tables* tps, float* grid, int B, int H, int W,
                         shineon_stream_t stream);

/* F.grid_sample(input, grid, mode="bilinear", padding_mode, align_corners=False)
 * input [B,C,Hin,Win] f32 NCHW, grid [B,Hout,Wout,2], out [B,C,Hout,Wout]. */
int shineon_grid_sample_fwd(const float* input, const float* grid, float* out, int B, int C, int Hin,
                            int Win, int Hout, int Wout, int padding_mode, shineon_stream_t stream);

/* Fused TPS + grid_sample: the grid is never materialised (warp_model.py:84-86 / 142-145).
 * Up to 3 inputs sampled with the same grid; inputs[i] is [B,C_i,H,W] f32 or NULL.
 * grid_out may be NULL. */
int shineon_tps_grid_sample_fwd(const float* theta, const shineon_tps_tables* tps, int B, int H, int W,
                                const float* in0, int C0, int pad0, float* out0,
                                const float* in1, int C1, int pad1, float* out1,
                                const float* in2, int C2, int pad2, float* out2,
                                float* grid_out, shineon_stream_t stream);

/* ------------------------------------------------------------------ */
/* N1: Resample2d (resample2d_kernel.cu:16-72 fwd, :76-198 bwd)         */
/* ------------------------------------------------------------------ */
/* in1 [B,C,Hi,Wi], flow [B,2,H,W] (pixel units), out [B,C,H,W]; all f32 NCHW. */
int shineon_resample2d_fwd(const float* in1, const float* flow, float* out, int B, int C, int Hi, int Wi,
                           int H, int W, int kernel_size, int bilinear, shineon_stream_t stream);
/* grad_in1 [B,C,Hi,Wi] must be zero-filled by the caller (resample2d.py:35); grad_flow [B,2,H,W]. */
int shineon_resample2d_bwd(const float* in1, const float* flow, const float* grad_out, float* grad_in1,
                           float* grad_flow, int B, int C, int Hi, int Wi, int H, int W, int kernel_size,
                           int bilinear, shineon_stream_t stream);

/* ------------------------------------------------------------------ */
/* N3: ChannelNorm (channelnorm_kernel.cu:19-60 fwd, :64-96 bwd)         */
/* ------------------------------------------------------------------ */
int shineon_channelnorm_fwd(const float* in, float* out, int B, int C, int H, int W, int norm_deg,
                            shineon_stream_t stream);
int shineon_channelnorm_bwd(const float* in, const float* out, const float* grad_out, float* grad_in,
                            int B, int C, int H, int W, int norm_deg, shineon_stream_t stream);

/* ------------------------------------------------------------------ */
/* N2: FlowNet Correlation (correlation_cuda_kernel.cu:47-147 fwd, :151-334 bwd) */
/* ------------------------------------------------------------------ */
/* Output geometry per correlation_cuda.cc:19-38. */
int shineon_correlation_out_shape(int C, int H, int W, int pad_size, int kernel_size, int max_displacement,
                                  int stride1, int stride2, int* out_c, int* out_h, int* out_w);
/* in1,in2 [B,C,H,W] f32 NCHW -> out [B,out_c,out_h,out_w].  No padded NHWC scratch (rbot1/2) needed. */
int shineon_correlation_fwd(const float* in1, const float* in2, float* out, int B, int C, int H, int W,
                            int pad_size, int kernel_size, int max_displacement, int stride1, int stride2,
                            shineon_stream_t stream);
/* Tensor-core path of the same op for kernel_size == 1, stride1 == 1: `full` f32 [B, H*W, H*W] holds the all-pairs
 * products <in1[p1,:], in2[p2,:]> (one tcgen05 GEMM per image, shineon_conv2d_igemm_fwd with w_per_image); this
 * kernel gathers the (2*(maxd/s2)+1)^2 displacements, scales by 1/C and zero-fills out-of-image taps. */
int shineon_correlation_gather(const float* full, float* out, int B, int C, int H, int W, int pad_size,
                               int max_displacement, int stride2, shineon_stream_t stream);
/* The same gather writing act(cost volume) as NHWC 16-bit planes [B,H,W,y_cstride] (y_hi / y_lo point at the first channel
 * of the consumer's channel window; channels D*D .. pad8(D*D) are written as zeros) -- FlowNetC.py:86-88: corr ->
 * corr_activation -> cat, without the NCHW f32 cost volume in between. */
int shineon_correlation_gather_planes(const float* full, void* y_hi, void* y_lo, int y_cstride, int B, int C, int H, int W,
                                      int pad_size, int max_displacement, int stride2, int act, float act_param,
                                      int plane_fmt, shineon_stream_t stream);
int shineon_correlation_bwd(const float* in1, const float* in2, const float* grad_out, float* grad_in1,
                            float* grad_in2, int B, int C, int H, int W, int pad_size, int kernel_size,
                            int max_displacement, int stride1, int stride2, shineon_stream_t stream);

/* ------------------------------------------------------------------ */
/* Tensor-core convolution (tcgen05 implicit GEMM)                      */
/* ------------------------------------------------------------------ */

/* OIHW f32 -> [Cout][kh][kw][cin_pad] bf16 hi (+ lo).  chan_map (device, int32[cin_pad]) maps a padded
 * input channel to the OIHW input channel or -1 (zero); NULL = identity for c<Cin, zero above.
 * transpose_io != 0 reads a ConvTranspose2d weight [Cin][Cout][kh][kw] with the taps flipped. */
int shineon_pack_conv_weight(const float* w, void* w_hi, void* w_lo, int Cout, int Cin, int kh, int kw,
                             int cin_pad, const int32_t* chan_map, int transpose_io, int plane_fmt,
                             float w_scale /* packed = w * w_scale; pass 1/w_scale as acc_scale */,
                             shineon_stream_t stream);
/* ConvTranspose2d(kernel 4, stride 2, padding 1) weight [Cin][Cout][4][4] (also: the data-gradient operand of a 4x4 s2
 * Conv2d, whose OIHW weight has exactly that layout with Cin := its Cout) -> the four phase-wise 2x2 stride-1 convs:
 * w_hi/w_lo [4 phases (py*2+px)][Cout][2*2][cin_pad].  submodules.py:34-38. */
int shineon_pack_deconv4x4s2_weight(const float* w, void* w_hi, void* w_lo, int Cin, int Cout, int cin_pad,
                                    const int32_t* chan_map, int plane_fmt, float w_scale, shineon_stream_t stream);

typedef struct shineon_conv2d_params {
  /* input activation, NHWC planes [N,H,W,cin_pad] bf16 */
  const void* x_hi;
  const void* x_lo; /* NULL => single 16-bit products (fast modes) */
  int N, H, W, cin_pad;
  int x_cstride; /* channels between consecutive pixels of x (0 = cin_pad); lets a conv read a 64-aligned channel
                    window of a wider concat buffer (the pointers already point at the window's first channel) */
  /* packed weights [Cout][kh*kw][cin_pad] bf16 */
  const void* w_hi;
  const void* w_lo; /* NULL unless x_lo given */
  int w_per_image;               /* != 0: weights are a batch [N][Cout][kh*kw][cin_pad], image n of x meets weight set n
                                    (per-image GEMMs, e.g. the FlowNetC cost volume: "weights" = the other feature map) */
  int Cout, kh, kw, stride;      /* stride 1 or 2 (stride 2 needs even H, W) */
  int pad_h, pad_w;              /* zero padding before the first row / column */
  int Ho, Wo;                    /* output size; rows/cols past the input read zeros, so any
                                    Ho <= (H + pad_h)/stride is legal (asymmetric padding) */
  /* epilogue: v = acc + bias; v = pre_act(v); v = v*scale + shift; v = post_act(v) */
  const float* bias;  /* [Cout] or NULL */
  const float* scale; /* [Cout] or NULL (with shift) */
  const float* shift;
  int pre_act, post_act;
  float act_param;
  float acc_scale; /* accumulator multiplier applied before the bias (0 = 1.0) */
  int plane_fmt;   /* encoding of x/w/y planes (shineon_plane_fmt) */
  /* outputs (any subset): f32 NHWC and/or bf16 planes NHWC, pixel (n, oh*oh_mul+oh_off, ow*ow_mul+ow_off)
   * of a [N,out_H,out_W,out_cstride] tensor, channels written at out_coffset.. */
  float* y_f32;
  void* y_hi;
  void* y_lo;
  int out_H, out_W, out_cstride, out_coffset;
  int oh_mul, oh_off, ow_mul, ow_off;
  /* tuning overrides, 0 = auto */
  int tile_n; /* 16,32,64,128,256 */
  int stages;
  /* optional: InstanceNorm statistics of the f32 output values, ACCUMULATED into [N][Cout][2] doubles (sum, sum of
   * squares; the caller zeroes them) for shineon_instnorm_act(stats_ready = 1) -- saves that call's statistics pass over
   * the output.  Needs y_f32 when the kernel's pixel tiles span several images (output smaller than 128 pixels). */
  double* stats_ws;
  /* K-blocks (64 input channels of one filter tap each) accumulated in one TMEM chain before the partial sum is taken
   * over in registers: 0 = default (16, i.e. 1024 K), < 0 = the whole K in one chain (tensor-core accumulation
   * truncates, so long chains lose accuracy: DESIGN.md section 4) */
  int acc_chunk_kb;
  /* optional split-K workspace (shineon_conv2d_igemm_fwd only): layers with fewer output tiles than half the SMs and a deep
   * K loop are cut along K so that tiles x K-slices cover the chip; every slice parks its partial sums here and the last
   * one to arrive adds them in slice order (bit-reproducible) and runs the epilogue.  Size from
   * shineon_conv2d_splitk_workspace_bytes() (0: this shape does not split); the first 256-byte-rounded total_tiles ints are
   * arrival counters that must be ZERO at the first launch (the kernel leaves them zero).  NULL = never split.  One
   * workspace per concurrently running launch. */
  void* splitk_ws;
  size_t splitk_ws_bytes;
  /* != 0: ConvTranspose2d(4, 2, 1) in ONE launch (shineon_conv2d_igemm_fwd only).  w_hi/w_lo = all four phase weight sets
   * [4 (py*2+px)][Cout][2*2][cin_pad] as shineon_pack_deconv4x4s2_weight writes them; kh = kw = 2, stride = 1, Ho = H,
   * Wo = W, out_H = 2H, out_W = 2W; pad_*, oh_*, ow_* are ignored: phase (py, px) pads (1-py, 1-px) and writes the output
   * pixels (2*oh + py, 2*ow + px).  Same results as four launches with oh_mul = ow_mul = 2 (submodules.py:34-38). */
  int deconv_phases;
} shineon_conv2d_params;
size_t shineon_conv2d_splitk_workspace_bytes(const shineon_conv2d_params* p);

int shineon_conv2d_igemm_fwd(const shineon_conv2d_params* p, shineon_stream_t stream);
/* Small-Cin first layers: the same GEMM with the im2col matrix produced INSIDE the kernel (producer warps gather the
 * NCHW f32 input(s) x0 [N,C0,H,W] (+ x1 [N,C1,H,W], concatenated on C), split to 16-bit planes and write the swizzled
 * shared-memory tile; nothing is materialised in HBM).  `p` describes the real convolution (N,H,W = input size,
 * kh,kw,stride,pad_*, Ho,Wo); w_hi/w_lo are packed tap-major: [Cout][1][cin_pad] with k = (fy*kw+fx)*C + c and
 * cin_pad = pad64(kh*kw*C); x_hi/x_lo are ignored. */
int shineon_conv2d_im2col_fwd(const shineon_conv2d_params* p, const float* x0, int C0, const float* x1, int C1,
                              shineon_stream_t stream);
/* CUDA-core fp32 direct convolution over the same operands: the on-GPU cross-check of the tcgen05
 * kernel used by tests (not on the product path). */
int shineon_conv2d_direct_fwd(const shineon_conv2d_params* p, shineon_stream_t stream);

/* ------------------------------------------------------------------ */
/* Layout / pointwise / normalisation                                   */
/* ------------------------------------------------------------------ */

/* NCHW f32 [N,C,H,W] (optionally a second tensor concatenated on C) -> NHWC planes [N,H,W,cpad],
 * with activation applied (unet.py:132 down-activation on the block input). */
int shineon_nchw_to_planes(const float* x0, int C0, const float* x1, int C1, void* y_hi, void* y_lo, int N,
                           int H, int W, int cpad, int act, float act_param, int plane_fmt,
                           shineon_stream_t stream);

/* Inverse: planes [N,H,W,(x_cstride)] (pointers at the first channel of the window) -> f32 NCHW [N,C,H,W]. */
int shineon_planes_to_nchw(const void* x_hi, const void* x_lo, int x_cstride, float* y, int N, int H, int W, int C,
                           int plane_fmt, shineon_stream_t stream);

/* Same, but written as the im2col matrix of a (kh x kw, stride, pad) convolution: y planes [N,Ho,Wo,kpad] with
 * k = (fy*kw+fx)*C + c.  Turns a small-Cin first layer (Cin 3/10/22) into a dense 1x1 GEMM. */
int shineon_nchw_im2col_planes(const float* x0, int C0, const float* x1, int C1, void* y_hi, void* y_lo, int N,
                               int H, int W, int kh, int kw, int stride, int pad, int Ho, int Wo, int kpad, int act,
                               float act_param, int plane_fmt, shineon_stream_t stream);

/* First layers with 4x4 / stride 2 / pad 1 filters and few input channels (GMM 22 ch warp.py:14, U-Net 10 ch
 * unet.py:129): shifted space-to-depth planes z [N, H/2+1, W/2+1, cpad] with z[Y][X][(py*2+px)*C + c] =
 * x[c][2Y-1+py][2X-1+px] (zero outside); the conv then is a 2x2 / stride 1 / pad 0 conv over z with the weight
 * re-indexed fy = 2*ay + py.  torch.cat([x0, x1], 1) is fused (x1 may be NULL). */
int shineon_nchw_s2d_planes(const float* x0, int C0, const float* x1, int C1, void* y_hi, void* y_lo, int N, int H, int W,
                            int cpad, int plane_fmt, shineon_stream_t stream);
/* col2im of a tap-stacked 3x3 convolution with few output channels: t f32 NHWC [N,H,W,tstride] holds the
 * 9*Cout per-input-pixel partial products (channel (fy*3+fx)*Cout+co); y f32 NHWC [N,H,W,Cout] = bias + shifted sum. */
int shineon_col2im3x3(const float* t, const float* bias, float* y, int N, int H, int W, int Cout, int tstride,
                      shineon_stream_t stream);

/* nn.Upsample(x2, bilinear, align_corners=False) -> Conv2d(3x3, pad 1) (unet.py:138-146,155,166) evaluated from
 * the tap-stacked products of the LOW-resolution tensor (the contraction over channels commutes with the
 * interpolation): t f32 NHWC [N,h,w,tstride], t[..., (fy*3+fx)*Cout + co] = <x[n,i,j,:], w[co,:,fy,fx]>
 * -> y f32 NHWC [N,2h,2w,Cout] = bias + sum over taps of the bilinearly upsampled t_tap, shifted by the tap,
 * zero where the tap leaves the 2h x 2w image (the conv's zero padding).
 * stats_ws (optional, like shineon_conv2d_params.stats_ws): [N][Cout][2] doubles accumulating sum / sum of squares of y. */
int shineon_upconv3x3_gather(const float* t, const float* bias, float* y, double* stats_ws, int N, int h, int w, int Cout,
                             int tstride, shineon_stream_t stream);

/* nn.InstanceNorm2d(affine=False, eps) over f32 NHWC x [N,H,W,C] (unet.py:133,135) followed by
 * activation; writes any subset of: f32 NHWC y_f32 (may alias x), planes y_hi/y_lo [N,H,W,cpad].
 * do_norm=0 skips the normalisation (innermost down block, unet.py:166-175).
 * stats_ws: caller-owned scratch of 2*N*C doubles (sum, sum of squares); zeroed and filled by the call, unless
 * stats_ready != 0: the producer of x accumulated them already (shineon_conv2d_params.stats_ws, shineon_upconv3x3_gather). */
int shineon_instnorm_act(const float* x, float* y_f32, void* y_hi, void* y_lo, double* stats_ws, int N, int H,
                         int W, int C, int cpad, float eps, int do_norm, int stats_ready, int act, float act_param,
                         int plane_fmt, shineon_stream_t stream);

/* Up-path input of a U-Net block (unet.py:138-146): up_act -> cat([skip, x'],C) -> bilinear x2
 * (align_corners=False).  Sources are activated planes [N,H,W,c{0,1}pad]; the extra activation
 * `act` (ReLU on the default path, where the skip already holds LeakyReLU'd values, unet.py:132,198)
 * is applied before interpolation.  Output planes [N,2H,2W,c0pad+c1pad]. src1 may be NULL. */
int shineon_upsample2x_cat(const void* s0_hi, const void* s0_lo, int c0pad, const void* s1_hi,
                           const void* s1_lo, int c1pad, void* y_hi, void* y_lo, int N, int H, int W,
                           int act, float act_param, int plane_fmt, shineon_stream_t stream);

/* SAGAN self-attention core (sagan.py:29-53) given the fused 1x1 projection
 * qkv f32 NHWC [N,HW,2*Cq+C] (q | k | v) and x f32 NHWC [N,HW,C]:
 * out = act(gamma * (V . softmax(q^T k)^T) + x); writes f32 and/or planes. */
int shineon_sagan_attention(const float* qkv, const float* x, const float* gamma, float* y_f32, void* y_hi,
                            void* y_lo, int N, int HW, int C, int Cq, int cpad, int act, float act_param,
                            int plane_fmt, shineon_stream_t stream);

/* ------------------------------------------------------------------ */
/* GMM glue                                                             */
/* ------------------------------------------------------------------ */

/* FeatureL2Norm x2 + FeatureCorrelation (warp.py:39-67) on f32 NHWC features [B,h*w,C]:
 * out planes [B,h,w,cpad] with channel iA = wA*h + hA (warp.py:60), value = <A/|A|, B/|B|>.
 * corr_f32 (optional) receives the same values as f32 NHWC [B,h,w,h*w].
 * normalize = 0: FeatureCorrelation.forward alone (warp.py:53-67) on features the caller normalised already. */
int shineon_l2norm_correlation(const float* featA, const float* featB, float* corr_f32, void* y_hi,
                               void* y_lo, int B, int h, int w, int C, int cpad, int plane_fmt, int normalize,
                               shineon_stream_t stream);
/* The same pair of FeatureL2Norms written as the operands of a per-image tensor-core GEMM (shineon_conv2d_igemm_fwd with
 * w_per_image): b_* [B,h,w,C] = scale * featB / |featB| (activation planes) and a_* [B][h*w][C] = scale * featA / |featA|
 * with rows in FeatureCorrelation's transposed pixel order iA = wA*h + hA (the per-image weights).  The 1x1 "conv" of b_*
 * with a_*, acc_scale = 1 / scale^2, then IS FeatureCorrelation's output [B,h,w,h*w].  C % 64 == 0; scale: a power of two. */
int shineon_l2norm_planes(const float* featA, const float* featB, void* a_hi, void* a_lo, void* b_hi, void* b_lo, int B, int h,
                          int w, int C, int plane_fmt, float scale, shineon_stream_t stream);
/* FeatureL2Norm.forward (warp.py:43-50) on the reference layout: y = x / sqrt(sum_c x^2 + 1e-6), f32 NCHW [B,C,H,W]. */
int shineon_feature_l2norm(const float* x, float* y, int B, int C, int H, int W, shineon_stream_t stream);

/* FeatureRegression tail (warp.py:94-99): x f32 NHWC [B,h,w,C] flattened in NCHW order ->
 * Linear(C*h*w -> out_dim) -> tanh.  weight [out_dim, C*h*w], bias [out_dim]; theta [B,out_dim]. */
int shineon_linear_tanh(const float* x, const float* weight, const float* bias, float* theta, int B, int h,
                        int w, int C, int out_dim, shineon_stream_t stream);

/* TPS warp of the decoded 8-bit cloth [B,H,W,3] (normalised on load like ToTensor + Normalize(0.5, 0.5)): the same
 * samples as shineon_tps_grid_sample_fwd(cloth f32, border) written (a) as f32 NCHW `warped` [B,3,H,W] and (b) split into
 * 16-bit hi/lo into the cloth channels of the U-Net stem's space-to-depth planes z [B,H/2+1,W/2+1,z_cstride]:
 * z[b][Y][X][(py*2+px)*ctot + c_off + c] = warped[b][c][2Y-1+py][2X-1+px]  (ctot channels per position). */
int shineon_tps_warp_u8_planes(const float* theta, const shineon_tps_tables* tps, const unsigned char* cloth_u8,
                               float* warped, void* z_hi, void* z_lo, int z_cstride, int ctot, int c_off, int plane_fmt,
                               int B, int H, int W, shineon_stream_t stream);

/* ------------------------------------------------------------------ */
/* U5: try-on composition (unet_mask_model.py:74-135)                   */
/* ------------------------------------------------------------------ */
/* unet_out f32 NHWC [B,H,W,Cout] (post InstanceNorm), Cout = (4 or 5)*n_frames.  Per frame f:
 *   rend = tanh(out[3f..3f+2]); mask = sigmoid(out[3n+f]); fmask = sigmoid(out[4n+f]) if flow_warp
 *   if warped_prev != NULL (frame f>0 with flows): rend' = (1-fmask)*warped_prev + fmask*rend
 *   tryon = (1-mask)*rend' + mask*cloth
 * Writes NCHW f32 channel slices of p_rendereds [B,3n,H,W], tryon_masks [B,n,H,W],
 * p_tryons [B,3n,H,W], flow_masks [B,n,H,W] for frame index f (any of them may be NULL = not wanted), and, when
 * p_tryons_u8 != NULL, the frame's try-on image as the reference saves it (shineon_image_to_u8) into
 * uint8 [B,n,H,W,3].  At least one of p_tryons / p_tryons_u8 must be given.
 * cloth is [B,3n,H,W]; warped_prev is [B,3,H,W] or NULL. */
int shineon_tom_compose(const float* unet_out, int Cout, const float* cloth, const float* warped_prev,
                        float* p_rendereds, float* tryon_masks, float* p_tryons, float* flow_masks,
                        unsigned char* p_tryons_u8, int B, int H, int W, int n_frames, int frame, int flow_warp,
                        shineon_stream_t stream);
/* U8 / 8f N2: the reference's image writer (visualization.py:73-76 save_images) without the PNG encoder:
 *   y = uint8(clamp((x + 1) * 0.5 * 255, 0, 255))  (truncation, like numpy astype), channel-last.
 * x f32 [B,C,H,W] -> y uint8 [B,H,W,C].  Bit-exact against the reference arithmetic on the same f32 input. */
int shineon_image_to_u8(const float* x, unsigned char* y, int B, int C, int H, int W, shineon_stream_t stream);

/* ------------------------------------------------------------------ */
/* F2: FlowNet2 glue (models/flownet2_pytorch/models.py:127-192, models/flownet.py:42-63) */
/* ------------------------------------------------------------------ */
/* inputs f32 [B,3,2,H,W] -> x f32 NCHW [B,6,H,W] = cat(frame0, frame1) of (inputs - mean_{f,h,w}) / rgb_max.
 * ws: 3*B doubles of scratch. */
int shineon_flownet_normalize(const float* inputs, float* x, double* ws, int B, int H, int W, float rgb_max,
                              shineon_stream_t stream);
/* nn.Upsample(scale_factor=4, bilinear|nearest) of the first two channels of src (f32 NHWC [B,h,w,src_cstride]),
 * times `mul`; dst f32 NCHW [B,2,4h,4w]. */
int shineon_upsample4x_flow(const float* src, int src_cstride, float* dst, int B, int h, int w, float mul,
                            int bilinear, shineon_stream_t stream);
/* `upsampled_flow*` = ConvTranspose2d(2, 2, 4, 2, 1) of a predicted flow (FlowNetC.py:59-62): flow f32 NHWC [B,h,w,flow_cstride]
 * (first two channels), weight f32 [2,2,4,4] (the module's own layout), bias [2] or NULL -> both channels of the consumer's
 * channel window, NHWC 16-bit planes [B,2h,2w,y_cstride] (y_hi / y_lo at the window's first channel). */
int shineon_flow_deconv4x4s2_planes(const float* flow, int flow_cstride, const float* weight, const float* bias, void* y_hi,
                                    void* y_lo, int y_cstride, int B, int h, int w, int plane_fmt, shineon_stream_t stream);
/* FlowNetS input: out [B,12,H,W] = [x | Resample2d(x[:,3:6], flow) | flow/div_flow | ChannelNorm(x[:,:3]-resampled)]. */
int shineon_flownet_warp_concat(const float* x, const float* flow, float* out, int B, int H, int W, float div_flow,
                                shineon_stream_t stream);
/* FlowNetFusion input: out [B,11,H,W] = [x[:,:3] | flow_sd | flow_s2 | |flow_sd| | |flow_s2| | diff_sd | diff_s2]. */
int shineon_flownet_fusion_concat(const float* x, const float* flow_sd, const float* flow_s2, float* out, int B, int H,
                                  int W, shineon_stream_t stream);
/* nn.Upsample(size=(Ho,Wo), mode="bilinear") (align_corners=False) times `mul`: FlowNet's pre/post resize for inputs
 * whose height is not a multiple of 64 (models/flownet.py:46-51,56-58).  x f32 [BC,Hi,Wi] -> y f32 [BC,Ho,Wo]. */
int shineon_bilinear_resize(const float* x, float* y, int BC, int Hi, int Wi, int Ho, int Wo, float mul,
                            shineon_stream_t stream);
/* conf [B,1,H,W] = (sum_c (im1 - Resample2d(im2, flow))^2 < threshold) as 0/1 floats. */
int shineon_flow_confidence(const float* im1, const float* im2, const float* flow, float* conf, int B, int C, int H,
                            int W, float threshold, shineon_stream_t stream);

/* ------------------------------------------------------------------ */
/* U6/U7 building block: fused Adam over flat f32 buffers (torch.optim.Adam semantics, base_model.py:165-168) */
/* ------------------------------------------------------------------ */
/* param -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps), with g = grad*grad_scale (+ weight_decay*param). */
int shineon_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n, float lr, float beta1,
                      float beta2, float eps, float weight_decay, int step, float grad_scale, shineon_stream_t stream);

/* ------------------------------------------------------------------ */
/* U6: backward of the conv layers (the reference gets these from cuDNN through autograd:               */
/*     models/unet_mask_model.py:137-217 training_step -> loss.backward() in the Lightning loop)         */
/* ------------------------------------------------------------------ */
/* Data gradients reuse shineon_conv2d_igemm_fwd on re-packed weights (shineon_pack_conv_weight with     */
/* transpose_io = 1: a stride-1 conv's dgrad is the flipped-tap conv with Cin/Cout swapped; a 4x4 s2 conv's */
/* dgrad is the four-phase 2x2 transposed conv).  Weight gradient:                                        */
/*   dW[co,(fy,fx),ci] = sum_{n,oh,ow} G[n,oh,ow,co] * X[n, oh*s+fy-p, ow*s+fx-p, ci]                      */
/* G, X are NHWC 16-bit planes (hi [+ lo]); tcgen05 with MN-major operands straight from the NHWC tiles. */
typedef struct shineon_conv2d_wgrad_params {
  const void* g_hi; /* grad of the conv output, planes [N,Ho,Wo,g_cstride], channels >= Cout zero */
  const void* g_lo; /* NULL in single-plane mode */
  int g_cpad;       /* channels the G boxes may touch (multiple of 64, >= Cout) */
  int g_cstride;    /* pixel pitch of G in elements (0 = g_cpad) */
  const void* x_hi; /* the conv's input planes [N,H,W,x_cstride] */
  const void* x_lo;
  int N, H, W, cin_pad, x_cstride; /* x_cstride 0 = cin_pad */
  int Cout, Cin;                   /* true channel counts of the parameter */
  int kh, kw, stride, pad_h, pad_w;
  int Ho, Wo;
  int plane_fmt;            /* shineon_plane_fmt of G and X */
  const int32_t* chan_map;  /* [cin_pad] packed channel -> weight input channel (-1 = padding), NULL = identity */
  int mode;                 /* 0: grad_w is OIHW; 1: im2col'd first layer (kh*kw taps folded into cin_pad, pass the
                               layer's true kh,kw,Cin and run with a 1x1 geometry); 2: ConvTranspose2d IOHW, flipped taps */
  float* grad_w;            /* the parameter's gradient (f32, the parameter's own layout) */
  float alpha, beta;        /* grad_w = beta*grad_w + alpha*dW  (alpha 0 = 1) */
  void* workspace;          /* >= shineon_conv2d_wgrad_workspace_bytes() */
  size_t workspace_bytes;
  int splits;               /* K-split override, 0 = auto (fills ~2 waves of 148 SMs) */
  int desc_variant;         /* 0 = default UMMA descriptor strides; 1 = swapped LBO/SBO (bring-up diagnostics) */
  int g_coffset;            /* first channel of G this parameter's outputs occupy (fused q|k|v projections); any value,
                               g_cpad then counts channels from g_coffset (rounded up to 64) */
} shineon_conv2d_wgrad_params;
size_t shineon_conv2d_wgrad_workspace_bytes(const shineon_conv2d_wgrad_params* p);
int shineon_conv2d_wgrad(const shineon_conv2d_wgrad_params* p, shineon_stream_t stream);
/* grad[c] = beta*grad[c] + alpha * sum over `pixels` rows of x[pixel*cstride + c]  (bias gradients; f64 accumulation).
 * workspace: C doubles. */
int shineon_channel_sum(const float* x, float* grad, void* workspace, long pixels, int C, int cstride, float alpha,
                        float beta, shineon_stream_t stream); /* x may point at a channel offset inside the row */

/* d(act o InstanceNorm2d): x = the conv output the forward normalised (f32 NHWC [N,H,W,C]), stats_fwd = the
 * forward's statistics workspace (shineon_instnorm_act), g1 (+ optional g2, summed) = dL/d(activated output).
 * Writes dL/dx as f32 NHWC and/or 16-bit planes [N,H,W,cpad].  do_norm = 0: plain g * act'(x).
 * stats_ws: 2*N*C doubles. */
int shineon_instnorm_act_bwd(const float* x, const double* stats_fwd, const float* g1, const float* g2, float* gx_f32,
                             void* gx_hi, void* gx_lo, double* stats_ws, int N, int H, int W, int C, int cpad, float eps,
                             int do_norm, int act, float act_param, int plane_fmt, shineon_stream_t stream);
/* gz[i] = (g1[i] + g2[i]) * act'(z[i])   (g2 may be NULL) */
int shineon_act_bwd(const float* z, const float* g1, const float* g2, float* gz, long n, int act, float act_param,
                    shineon_stream_t stream);
/* Adjoint of shineon_upsample2x_cat: g_up f32 NHWC [N,2H,2W,g_cstride] (channels [0,C0) = first source, [C0,C0+C1) =
 * second) -> g0 [N,H,W,C0], g1 [N,H,W,C1] (NULL when C1 == 0). */
int shineon_upsample2x_cat_bwd(const float* g_up, int g_cstride, float* g0, int C0, float* g1, int C1, int N, int H, int W,
                               shineon_stream_t stream);
/* SAGAN attention backward (attention/sagan.py:29-53): qkv f32 [N,HW,2Cq+C] (the forward's projections), g_out =
 * dL/d(gamma*o + x) f32 [N,HW,C] -> g_qkv f32 [N,HW,2Cq+C], g_gamma[0] = beta_gamma*g_gamma[0] + dL/dgamma.
 * The residual branch (dL/dx += g_out) and the 1x1 projections' dgrad/wgrad are the caller's. */
size_t shineon_sagan_attention_bwd_workspace_bytes(int N, int HW);
int shineon_sagan_attention_bwd(const float* qkv, const float* gamma, const float* g_out, float* g_qkv, float* g_gamma,
                                void* workspace, size_t workspace_bytes, int N, int HW, int C, int Cq, float beta_gamma,
                                shineon_stream_t stream);
/* Backward of shineon_tom_compose for one frame: gradients wrt THIS frame's outputs as per-frame NCHW tensors
 * (g_rendereds / g_tryons [B,3,H,W], g_masks / g_flow_masks [B,1,H,W]; any may be NULL = zero) ->
 * g_unet_out f32 NHWC (only this frame's channels are written) and, with flow-warp, g_warped_prev [B,3,H,W]. */
int shineon_tom_compose_bwd(const float* unet_out, int Cout, const float* cloth, const float* warped_prev,
                            const float* g_rendereds, const float* g_masks, const float* g_tryons,
                            const float* g_flow_masks, float* g_unet_out, float* g_warped_prev, int B, int H, int W,
                            int n_frames, int frame, int flow_warp, shineon_stream_t stream);
/* loss[0] = beta_loss*loss[0] + loss_weight*mean|a-b|; grad_a (may be NULL) (+)= loss_weight*sign(a-b)/n
 * (F.l1_loss, unet_mask_model.py:174-188; loss.py:106-122).  workspace: one double. */
int shineon_l1_loss(const float* a, const float* b, float* grad_a, float* loss, void* workspace, long n, float loss_weight,
                    float beta_loss, int accumulate_grad, shineon_stream_t stream);
/* VGG19 glue (models/networks/vgg.py:6-38): 2x2/2 max-pool on f32 NHWC -> f32 and/or planes; backward routes to the
 * first maximum of the window in scan order (ATen). */
int shineon_maxpool2x2_fwd(const float* x, float* y_f32, void* y_hi, void* y_lo, int N, int H, int W, int C, int cpad,
                           int plane_fmt, shineon_stream_t stream);
/* y [N,C,H,W] (+)= x [N,H,W,x_cstride][..., :C]  (image gradients produced by NHWC kernels back to the reference layout) */
int shineon_nhwc_to_nchw_add(const float* x, int x_cstride, float* y, int N, int H, int W, int C, int accumulate,
                             shineon_stream_t stream);
int shineon_maxpool2x2_bwd(const float* x, const float* g_y, int g_cstride, float* g_x, int N, int H, int W, int C,
                           shineon_stream_t stream);

/* ------------------------------------------------------------------ */
/* Dataset-side per-frame tensor prep                    */
/* ------------------------------------------------------------------ */

/* Pillow's 8-bit BILINEAR resampling coefficients (Resample.c precompute_coeffs + normalize_coeffs_8bpc) for one axis,
 * computed on the HOST: bounds[2*out] = (first input index, tap count), kk[out*ksize] = 22-bit fixed-point weights.
 * Returns ksize (call with kk = NULL to query it) or a negative status.  Replaces the two
 * Image.resize(..., Image.BILINEAR) calls of get_person_body_silhouette (datasets/tryon_dataset.py:352-358). */
int shineon_pil_bilinear_coeffs(int in_size, int out_size, int* bounds, int* kk);

/* One call = TryonDataset.get_person_representation + get_cloth_representation (datasets/tryon_dataset.py:156-175,
 * 203-251, 323-447) for F frames from decoded 8-bit images (channel-last, as np.array(PIL.Image) gives them).
 * Any input / output pair may be NULL.  Outputs are f32 NCHW device tensors, bit-identical to the reference's. */
typedef struct {
  const unsigned char* image;     /* [F,H,W,3] person frame                                  */
  const unsigned char* parse;     /* [F,H,W]   LIP label map                                 */
  const unsigned char* cloth;     /* [F,H,W,3]                                               */
  const unsigned char* densepose; /* [F,H,W,3]                                               */
  const double* pose;             /* [F,n_joints,3] (x, y, confidence); only for im_cocopose */
  float* image_out;               /* [F,3,H,W]  ToTensor + Normalize(0.5, 0.5)               */
  float* cloth_out;               /* [F,3,H,W]                                               */
  float* cloth_mask_out;          /* [F,1,H,W]  where(cloth >= threshold, 0, 1)[0]           */
  float* densepose_out;           /* [F,3,H,W]                                               */
  float* agnostic_out;            /* [F,4,H,W]  silhouette | head                            */
  float* cocopose_out;            /* [F,n_joints,H,W] constant -1 (tryon_dataset.py:415-423) */
  float* im_cocopose_out;         /* [F,1,H,W]  joint squares, -1 / 1                        */
  const int* tab_bounds[4];       /* device copies of shineon_pil_bilinear_coeffs tables for */
  const int* tab_kk[4];           /*   W -> W/16, H -> H/16, W/16 -> W, H/16 -> H            */
  int tab_ksize[4];
  int F, H, W, n_joints, radius;
  float cloth_mask_threshold;
} shineon_frame_prep_params;
int shineon_frame_prep(const shineon_frame_prep_params* p, shineon_stream_t stream);

/* The same prep written straight into the first layers' operands (TryOnPipeline.run_raw): instead of f32 NCHW tensors
 * (which torch.cat + shineon_nchw_s2d_planes / shineon_nchw_im2col_planes would re-read and re-write) it emits 16-bit
 * hi (+ lo) planes, bit-identical to that chain:
 *   gmm   [F,H/2+1,W/2+1,gmm_cpad]   shifted space-to-depth of cat(agnostic, cocopose)            (WarpModel person input)
 *   unet  [F,H/2+1,W/2+1,unet_cpad]  the same of cat(agnostic, densepose, cloth'): cloth' (channels 7..9 of each of the
 *                                    four positions) is left zero for shineon_tps_warp_u8_planes   (U-Net stem input)
 *   cloth [F,H/2,W/2,cloth_cpad]     im2col (4x4, stride 2, pad 1; k = (fy*4+fx)*3 + c) of the cloth (extractionB stem)
 * silhouette_scratch: F*H*W bytes (the 8-bit silhouette after Pillow's two resizes). */
typedef struct {
  const unsigned char* image;     /* [F,H,W,3] */
  const unsigned char* parse;     /* [F,H,W]   */
  const unsigned char* cloth;     /* [F,H,W,3] */
  const unsigned char* densepose; /* [F,H,W,3] */
  unsigned char* silhouette_scratch;
  void *gmm_hi, *gmm_lo, *unet_hi, *unet_lo, *cloth_hi, *cloth_lo; /* lo: all NULL in single-plane modes */
  int gmm_cpad, unet_cpad, cloth_cpad;
  int plane_fmt;
  const int* tab_bounds[4];       /* as in shineon_frame_prep_params */
  const int* tab_kk[4];
  int tab_ksize[4];
  int F, H, W, n_joints;
} shineon_frame_prep_planes_params;
int shineon_frame_prep_planes(const shineon_frame_prep_planes_params* p, shineon_stream_t stream);

/* Middlebury .flo payload (the interleaved f32 u,v after the 12-byte header) -> f32 [2,H,W] = (x - 0.5) / 0.5
 * (flownet2_pytorch/utils/flow_utils.py:7-26 + flow_norm, datasets/tryon_dataset.py:121,288-289). */
int shineon_flo_decode(const void* flo_payload, float* out, int H, int W, shineon_stream_t stream);

/* ------------------------------------------------------------------ */
/* SAMS generator: SPADE modulation and the nearest     */
/* resizes around the blocks                                           */
/* ------------------------------------------------------------------ */

/* Per-(image, channel) sum and sum of squares of f32 NHWC x [N,HW,C] into stats_ws (2*N*C doubles, zeroed by the call):
 * the statistics nn.InstanceNorm2d(affine=False) needs (SPADE's param_free_norm, models/networks/sams/spade.py:55,71). */
int shineon_chan_stats(const float* x, double* stats_ws, int N, int HW, int C, shineon_stream_t stream);

/* SPADE.forward's modulation (models/networks/sams/spade.py:68-84) on f32 NHWC x [N,H,W,C]:
 *   y = act( norm(x) * gb[..., c] + gb[..., C + c] )
 * norm_mode 0: identity; 1: InstanceNorm2d from stats_ws (shineon_chan_stats); 2: eval-mode BatchNorm2d /
 * SynchronizedBatchNorm2d(affine=False) as per-channel nscale / nshift [C].  gb: f32 NHWC [N,H,W,gb_cstride], the output
 * of ONE conv holding (1 + gamma) in channels [0,C) and beta in [C,2C) (mlp_gamma / mlp_beta stacked, the "+1" folded into
 * the bias).  act = the activation AnySpadeResBlock applies next (spade.py:157-158), 0 for norm_s.
 * Writes any subset of f32 y_f32 (row stride y_cstride; may be a channel window) and planes y_hi / y_lo (stride cpad). */
int shineon_spade_modulate(const float* x, const double* stats_ws, const float* nscale, const float* nshift,
                           const float* gb, int gb_cstride, float* y_f32, int y_cstride, void* y_hi, void* y_lo, int N,
                           int H, int W, int C, int cpad, float eps, int norm_mode, int act, float act_param,
                           int plane_fmt, shineon_stream_t stream);

/* nn.Upsample(scale_factor=s) (mode nearest; sams_generator.py:295-310) on f32 NHWC: y[n,h,w,:] = x[n, min(floor(h *
 * scale_h), Hs-1), min(floor(w * scale_w), Ws-1), :] with scale = 1/s (PyTorch's source-index rule). */
int shineon_nearest_resize_nhwc(const float* x, float* y, int N, int Hs, int Ws, int H, int W, int C, float scale_h,
                                float scale_w, shineon_stream_t stream);

/* F.interpolate(segmap, size=(H,W), mode="nearest") (spade.py:74) of an f32 NCHW label map [N,C,Hs,Ws], written as the
 * 16-bit planes [N,H,W,cpad] the mlp_shared conv reads (padding channels zero).  scale = Hs / H, Ws / W. */
int shineon_nearest_resize_planes(const float* x, int N, int C, int Hs, int Ws, void* y_hi, void* y_lo, int H, int W,
                                  int cpad, float scale_h, float scale_w, int plane_fmt, shineon_stream_t stream);

/* The same resize fused with the im2col of the conv that reads it (mlp_shared: ks x ks, stride 1, pad ks/2, spade.py:60-62):
 * planes [N,H,W,kpad], k = (fy*ks + fx)*C + c, zero outside the resized map and for k >= ks*ks*C.  The conv is then a dense
 * 1x1 GEMM over K = kpad with the weight reordered tap-major (label maps have 2-18 channels). */
int shineon_nearest_im2col_planes(const float* x, int N, int C, int Hs, int Ws, void* y_hi, void* y_lo, int H, int W, int ks,
                                  int kpad, float scale_h, float scale_w, int plane_fmt, shineon_stream_t stream);

/* y = a + b over n f32 elements (the residual x_s + dx of AnySpadeResBlock.forward, spade.py:160); y may alias a or b. */
int shineon_add_nhwc(const float* a, const float* b, float* y, long n, shineon_stream_t stream);

/* The tail of SamsModel.generate_n_frames (models/sams_model.py:226-236): gen_out f32 NHWC [B,H,W,Cg] = frame (3) [+ blend
 * weight (1)]; out (one frame slot of the [b,n,3,h,w] buffer, batch stride out_bstride floats, NCHW) =
 * (1 - w) * warped_prev + w * frame when warped_prev (f32 NCHW [B,3,H,W], the Resample2d of the last frame) is given
 * (Cg == 4), else the frame (Cg == 3). */
int shineon_sams_flow_blend(const float* gen_out, int Cg, const float* warped_prev, float* out, long out_bstride, int B,
                            int H, int W, shineon_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SHINEON_B200_H_ */
