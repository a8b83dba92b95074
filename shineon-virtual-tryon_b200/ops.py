"""Torch-tensor wrappers over the C ABI: allocation, argument checks and stream plumbing only.

All compute happens in libshineon_b200.so; nothing here falls back to torch ops.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ACT, PAD, Conv2dParams, TpsTables, check


# When set to a list, conv2d() brackets every tensor-core conv launch with CUDA events on the launching stream
# and appends (algorithmic_flops, start_event, end_event); bench.py uses it for the roofline of the dominant kernel.
PROFILE = None
# K-blocks per TMEM accumulation chain of the conv kernel (shineon_conv2d_params.acc_chunk_kb): 0 = kernel default (16),
# -1 = one chain over the whole K (the round-1 behaviour; kept for the accuracy A/B in tests/diag_accum.py)
ACC_CHUNK_KB = 0
# split-K for layers with fewer output tiles than half the SMs and a deep K loop (shineon_conv2d_params.splitk_ws).
# OFF by default: measured on B200 as graph replays (tests/prof_splitk.py, profiles/r02_splitk.md) it wins 10-17 % on the
# 16-tile layers (8x6 512->256 at 80 frames: 75 -> 67 us; FlowNet 4x3 1024->1024: 82 -> 68 us) and loses up to 2x wherever
# tiles x slices exceeds the SM count; the whole try-on step got 1.7 % slower, FlowNet2 8 %.  Kept (and tested) as an option.
SPLIT_K = False


def C_void(v):
    return C.c_void_p(v)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _req(t, dtype=torch.float32, name="tensor"):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise _lib.ShineonError(f"{name}: expected a CUDA tensor (shineon ops have no CPU path)")
    if t.dtype != dtype:
        raise _lib.ShineonError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.ShineonError(f"{name}: expected a contiguous tensor")
    return t


def cpad64(c):
    return (c + 63) // 64 * 64


# Numeric modes of the tensor-core path: name -> (plane format, hi/lo split)
PRECISIONS = {
    "fp16x3": (_lib.FMT_FP16, True),   # default: 22 mantissa bits per operand, 3 MMAs per K-slice (fp32-grade)
    "bf16x3": (_lib.FMT_BF16, True),   # range-safe variant: 16 mantissa bits per operand
    "fp16": (_lib.FMT_FP16, False),    # single products (fast modes)
    "bf16": (_lib.FMT_BF16, False),
}
DEFAULT_PRECISION = "fp16x3"
# Inference U-Net decoders evaluate upsample -> conv3x3 at the low resolution (UpsampledConv3x3); False = materialise
# the upsampled concat (upsample2x_cat) and convolve at the high resolution, as the training tape does.
UPCONV_LOWRES = True


def resolve_precision(p):
    """Accepts a mode name, a (fmt, split) pair, or the legacy bool (True = default split mode, False = bf16)."""
    if p is None or p is True:
        p = DEFAULT_PRECISION
    elif p is False:
        p = "bf16"
    if isinstance(p, str):
        return PRECISIONS[p]
    return p


_DTYPES = {_lib.FMT_BF16: torch.bfloat16, _lib.FMT_FP16: torch.float16}


class Planes:
    """NHWC activation as 16-bit hi (+ lo) planes [N,H,W,cpad]; x ~= hi + lo."""

    def __init__(self, N, H, W, C, prec=None, device="cuda", cpad=None, zero_pad=True):
        self.fmt, split = resolve_precision(prec)
        self.N, self.H, self.W, self.C = N, H, W, C
        self.cpad = cpad64(C) if cpad is None else cpad
        # padding channels must hold zeros (they meet zero weights, but NaN * 0 = NaN); producers that write the
        # padding themselves pass zero_pad=False
        alloc = torch.empty if (self.cpad == C or not zero_pad) else torch.zeros
        self.hi = alloc(N, H, W, self.cpad, dtype=_DTYPES[self.fmt], device=device)
        self.lo = alloc(N, H, W, self.cpad, dtype=_DTYPES[self.fmt], device=device) if split else None

    @property
    def prec(self):
        return (self.fmt, self.lo is not None)

    # a Planes may be a 64-aligned channel window of a wider (concat) buffer
    coffset = 0

    @property
    def cstride(self):
        return self.hi.shape[-1]

    def window(self, c0, C, align=64):
        """View of channels [c0, c0+C) sharing this buffer's storage.  align = 64 (default): the window starts on a K-block
        and owns whole K-blocks, so it can also be read as a conv input on its own; align = 8: a write-only slot of a packed
        concat buffer (16-byte aligned stores), whose neighbours start right behind its 8-channel padding."""
        assert align in (8, 64) and c0 % align == 0
        wpad = cpad64(C) if align == 64 else (C + 7) // 8 * 8
        assert c0 + wpad <= self.hi.shape[-1]
        v = Planes.__new__(Planes)
        v.fmt, v.N, v.H, v.W, v.C, v.cpad = self.fmt, self.N, self.H, self.W, C, wpad
        v.hi, v.lo, v.coffset = self.hi, self.lo, self.coffset + c0
        return v

    def _ptr(self, t):
        return C_void(0 if t is None else t.data_ptr() + 2 * self.coffset)

    def float(self):
        """Reconstructed f32 NCHW tensor (debug / tests)."""
        v = self.hi.float()
        if self.lo is not None:
            v = v + self.lo.float()
        return v[..., self.coffset: self.coffset + self.C].permute(0, 3, 1, 2).contiguous()


# ----------------------------------------------------------------------------- TPS / grid sample
class TpsTablesDev:
    """Device copies of TpsGridGen's constants (built by the host module like warp.py:116-157)."""

    def __init__(self, Li, P_X, P_Y, grid_X, grid_Y, grid_size, device):
        f = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous().view(-1)
        self.Li, self.P_X, self.P_Y, self.grid_X, self.grid_Y = f(Li), f(P_X), f(P_Y), f(grid_X), f(grid_Y)
        self.grid_size = int(grid_size)
        self.struct = TpsTables(_p(self.Li), _p(self.P_X), _p(self.P_Y), _p(self.grid_X), _p(self.grid_Y),
                                self.grid_size)


def tps_grid(theta, tables, H, W):
    theta = _req(theta, name="theta")
    B = theta.shape[0]
    grid = torch.empty(B, H, W, 2, dtype=torch.float32, device=theta.device)
    check(_lib.load().shineon_tps_grid_fwd(_p(theta), C.byref(tables.struct), _p(grid), B, H, W, _stream()),
          "shineon_tps_grid_fwd")
    return grid


def grid_sample(inp, grid, padding_mode="zeros"):
    inp, grid = _req(inp, name="input"), _req(grid, name="grid")
    B, Cc, Hin, Win = inp.shape
    _, Hout, Wout, _ = grid.shape
    out = torch.empty(B, Cc, Hout, Wout, dtype=torch.float32, device=inp.device)
    check(_lib.load().shineon_grid_sample_fwd(_p(inp), _p(grid), _p(out), B, Cc, Hin, Win, Hout, Wout,
                                              PAD[padding_mode], _stream()), "shineon_grid_sample_fwd")
    return out


def tps_grid_sample(theta, tables, H, W, inputs, want_grid=False):
    """inputs: list of up to 3 (tensor [B,C,H,W], padding_mode).  Returns (outs, grid or None)."""
    theta = _req(theta, name="theta")
    B = theta.shape[0]
    args, outs = [], []
    for i in range(3):
        if i < len(inputs):
            t, mode = inputs[i]
            t = _req(t, name=f"input{i}")
            assert t.shape[0] == B and t.shape[2] == H and t.shape[3] == W
            o = torch.empty_like(t)
            outs.append(o)
            args += [_p(t), t.shape[1], PAD[mode], _p(o)]
        else:
            args += [_p(None), 0, 0, _p(None)]
    grid = torch.empty(B, H, W, 2, dtype=torch.float32, device=theta.device) if want_grid else None
    check(_lib.load().shineon_tps_grid_sample_fwd(_p(theta), C.byref(tables.struct), B, H, W, *args, _p(grid),
                                                  _stream()), "shineon_tps_grid_sample_fwd")
    return outs, grid


# ----------------------------------------------------------------------------- FlowNet2 native ops
def resample2d_fwd(in1, flow, kernel_size=1, bilinear=True):
    in1, flow = _req(in1, name="input1"), _req(flow, name="input2")
    _, d, Hi, Wi = in1.shape
    b, _, h, w = flow.shape
    out = torch.empty(b, d, h, w, dtype=torch.float32, device=in1.device)
    check(_lib.load().shineon_resample2d_fwd(_p(in1), _p(flow), _p(out), b, d, Hi, Wi, h, w, kernel_size,
                                             int(bool(bilinear)), _stream()), "shineon_resample2d_fwd")
    return out


def resample2d_bwd(in1, flow, grad_out, kernel_size=1, bilinear=True, grad_in1=None):
    """grad_in1: optional existing accumulator (the kernel scatters with atomicAdd, so it adds onto its content)."""
    in1, flow, grad_out = _req(in1), _req(flow), _req(grad_out.contiguous())
    _, d, Hi, Wi = in1.shape
    b, _, h, w = flow.shape
    g1 = torch.zeros_like(in1) if grad_in1 is None else _req(grad_in1, name="grad_in1")
    g2 = torch.empty_like(flow)
    check(_lib.load().shineon_resample2d_bwd(_p(in1), _p(flow), _p(grad_out), _p(g1), _p(g2), b, d, Hi, Wi, h, w,
                                             kernel_size, int(bool(bilinear)), _stream()), "shineon_resample2d_bwd")
    return g1, g2


def channelnorm_fwd(x, norm_deg=2):
    x = _req(x)
    b, c, h, w = x.shape
    out = torch.empty(b, 1, h, w, dtype=torch.float32, device=x.device)
    check(_lib.load().shineon_channelnorm_fwd(_p(x), _p(out), b, c, h, w, norm_deg, _stream()),
          "shineon_channelnorm_fwd")
    return out


def channelnorm_bwd(x, out, grad_out, norm_deg=2):
    x, out, grad_out = _req(x), _req(out), _req(grad_out.contiguous())
    b, c, h, w = x.shape
    g = torch.empty_like(x)
    check(_lib.load().shineon_channelnorm_bwd(_p(x), _p(out), _p(grad_out), _p(g), b, c, h, w, norm_deg, _stream()),
          "shineon_channelnorm_bwd")
    return g


def correlation_out_shape(Cc, H, W, pad_size, kernel_size, max_displacement, stride1, stride2):
    oc, oh, ow = C.c_int(), C.c_int(), C.c_int()
    check(_lib.load().shineon_correlation_out_shape(Cc, H, W, pad_size, kernel_size, max_displacement, stride1,
                                                    stride2, C.byref(oc), C.byref(oh), C.byref(ow)),
          "shineon_correlation_out_shape")
    return oc.value, oh.value, ow.value


def correlation_fwd(in1, in2, pad_size, kernel_size, max_displacement, stride1, stride2, tensor_cores=True):
    in1, in2 = _req(in1), _req(in2)
    b, c, h, w = in1.shape
    if kernel_size == 1 and stride1 == 1 and pad_size == max_displacement and tensor_cores:
        return correlation_planes(nchw_to_planes(in1), nchw_to_planes(in2), c, pad_size, max_displacement, stride2)
    oc, oh, ow = correlation_out_shape(c, h, w, pad_size, kernel_size, max_displacement, stride1, stride2)
    out = torch.empty(b, oc, oh, ow, dtype=torch.float32, device=in1.device)
    check(_lib.load().shineon_correlation_fwd(_p(in1), _p(in2), _p(out), b, c, h, w, pad_size, kernel_size,
                                              max_displacement, stride1, stride2, _stream()),
          "shineon_correlation_fwd")
    return out


def correlation_planes(f1, f2, C, pad_size, max_displacement, stride2, out_planes=None, act=None, act_param=0.0):
    """Tensor-core cost volume (kernel_size 1, stride1 1, pad == max_displacement): f1, f2 are Planes [B,H,W,C]
    (e.g. straight out of the conv3 layers).  One per-image tcgen05 GEMM (f2 plays the weights) + a gather.
    Returns f32 NCHW [B, D*D, H, W]; with out_planes (a Planes / channel window of >= D*D channels) the gather writes
    act(cost volume) there as NHWC planes instead and out_planes is returned."""
    assert f1.prec == f2.prec and f1.cpad == f2.cpad and f1.cstride == f1.cpad and f2.cstride == f2.cpad
    B, H, W = f1.N, f1.H, f1.W
    P = H * W
    pc = PackedConv.__new__(PackedConv)
    pc.fmt = f2.fmt
    pc.Cout, pc.Cin, pc.kh, pc.kw, pc.stride, pc.pad_h, pc.pad_w = P, C, 1, 1, 1, 0, 0
    pc.cin_pad, pc.w_hi, pc.w_lo, pc.bias, pc.acc_scale, pc.transposed, pc.per_image = f2.cpad, f2.hi, f2.lo, None, 1.0, False, True
    full, _ = conv2d(f1, pc, want_f32=True)  # [B,H,W,P]
    D = 2 * (max_displacement // stride2) + 1
    if out_planes is not None:
        assert out_planes.N == B and out_planes.H == H and out_planes.W == W and out_planes.C >= D * D and out_planes.fmt == f1.fmt
        check(_lib.load().shineon_correlation_gather_planes(_p(full), out_planes._ptr(out_planes.hi), out_planes._ptr(out_planes.lo),
                                                            out_planes.cstride, B, C, H, W, pad_size, max_displacement, stride2,
                                                            ACT[act], float(act_param), out_planes.fmt, _stream()),
              "shineon_correlation_gather_planes")
        return out_planes
    out = torch.empty(B, D * D, H, W, dtype=torch.float32, device=full.device)
    check(_lib.load().shineon_correlation_gather(_p(full), _p(out), B, C, H, W, pad_size, max_displacement, stride2,
                                                 _stream()), "shineon_correlation_gather")
    return out


def correlation_bwd(in1, in2, grad_out, pad_size, kernel_size, max_displacement, stride1, stride2):
    in1, in2, grad_out = _req(in1), _req(in2), _req(grad_out.contiguous())
    b, c, h, w = in1.shape
    g1, g2 = torch.empty_like(in1), torch.empty_like(in2)
    check(_lib.load().shineon_correlation_bwd(_p(in1), _p(in2), _p(grad_out), _p(g1), _p(g2), b, c, h, w, pad_size,
                                              kernel_size, max_displacement, stride1, stride2, _stream()),
          "shineon_correlation_bwd")
    return g1, g2


# ----------------------------------------------------------------------------- tensor-core conv
class PackedConv:
    """Conv2d / ConvTranspose2d weight repacked to [Cout][kh*kw][cin_pad] bf16 hi/lo on the device."""

    def __init__(self, weight, bias=None, stride=1, pad=0, cin_pad=None, prec=None, chan_map=None,
                 transposed=False, pad_hw=None):
        self.fmt, split = resolve_precision(prec)
        weight = _req(weight.detach().float().contiguous(), name="weight")
        if transposed:
            Cin, Cout, kh, kw = weight.shape
        else:
            Cout, Cin, kh, kw = weight.shape
        self.Cout, self.Cin, self.kh, self.kw, self.stride = Cout, Cin, kh, kw, stride
        self.pad_h, self.pad_w = pad_hw if pad_hw is not None else (pad, pad)
        if cin_pad is None:
            cin_pad = cpad64(Cin)
        self.cin_pad = cin_pad
        dev = weight.device
        self.w_hi = torch.empty(Cout, kh * kw, cin_pad, dtype=_DTYPES[self.fmt], device=dev)
        self.w_lo = torch.empty_like(self.w_hi) if split else None
        # fp16 planes: scale the weights by a power of two so hi/lo stay out of the fp16 subnormal range; the
        # conv epilogue multiplies the accumulator by the exact inverse (acc_scale).
        w_scale = 1.0
        if self.fmt == _lib.FMT_FP16:
            amax = float(weight.abs().max())
            if amax > 0:
                import math

                w_scale = 2.0 ** math.floor(math.log2(16384.0 / amax))
        self.acc_scale = 1.0 / w_scale
        cm = None
        if chan_map is not None:
            cm = torch.as_tensor(chan_map, dtype=torch.int32, device=dev).contiguous()
            assert cm.numel() == cin_pad
        check(_lib.load().shineon_pack_conv_weight(_p(weight), _p(self.w_hi), _p(self.w_lo), Cout, Cin, kh, kw,
                                                   cin_pad, _p(cm), int(transposed), self.fmt, w_scale, _stream()),
              "shineon_pack_conv_weight")
        self.bias = None if bias is None else _req(bias.detach().float().contiguous(), name="bias")
        self.transposed = transposed


def conv2d(x, pc, *, scale=None, shift=None, pre_act=None, post_act=None, act_param=0.0, out_f32=None,
           out_planes=None, out_coffset=0, want_f32=False, want_planes=False, direct=False, tile_n=0, stages=0,
           out_geom=None, out_hw=None, stats_ws=None, deconv_phases=False):
    """Run one convolution layer on planes `x` with packed weights `pc`.

    Outputs: f32 NHWC tensor [N,Ho,Wo,Cout] (want_f32 / out_f32) and/or Planes (want_planes / out_planes).
    out_geom = (out_H, out_W, oh_mul, oh_off, ow_mul, ow_off) scatters into a larger output (deconv phases).
    deconv_phases: `pc` holds the four phase weight sets of a ConvTranspose2d(4, 2, 1) ([4, Cout, 4, cin_pad]); one launch
    computes all four output parities (networks/deconv.py).
    stats_ws: zeroed f64 [N*Cout*2] buffer; the kernel adds the per-(image, channel) sum / sum of squares of the f32
    output values to it (InstanceNorm statistics for instnorm_act(stats_ready=True)).
    """
    N, H, W = x.N, x.H, x.W
    assert x.cpad == pc.cin_pad, f"activation cpad {x.cpad} != packed cin_pad {pc.cin_pad}"
    assert (x.lo is None) == (pc.w_lo is None) and x.fmt == pc.fmt, "precision of activation and weights must agree"
    if out_hw is not None:
        Ho, Wo = out_hw
    else:
        Ho = (H + 2 * pc.pad_h - pc.kh) // pc.stride + 1
        Wo = (W + 2 * pc.pad_w - pc.kw) // pc.stride + 1
    oH, oW, ohm, oho, owm, owo = out_geom if out_geom else (Ho, Wo, 1, 0, 1, 0)
    dev = x.hi.device
    if want_f32 and out_f32 is None:
        out_f32 = torch.empty(N, oH, oW, pc.Cout, dtype=torch.float32, device=dev)
    if want_planes and out_planes is None:
        out_planes = Planes(N, oH, oW, pc.Cout, prec=x.prec, device=dev)
    p = Conv2dParams()
    p.x_hi, p.x_lo = x._ptr(x.hi), x._ptr(x.lo)
    p.N, p.H, p.W, p.cin_pad, p.x_cstride = N, H, W, x.cpad, x.cstride
    p.w_hi, p.w_lo = _p(pc.w_hi), _p(pc.w_lo)
    p.w_per_image = int(getattr(pc, "per_image", False))
    p.Cout, p.kh, p.kw, p.stride, p.pad_h, p.pad_w = pc.Cout, pc.kh, pc.kw, pc.stride, pc.pad_h, pc.pad_w
    p.Ho, p.Wo = Ho, Wo
    p.bias, p.scale, p.shift = _p(pc.bias), _p(scale), _p(shift)
    p.pre_act, p.post_act, p.act_param = ACT[pre_act], ACT[post_act], float(act_param)
    p.acc_scale, p.plane_fmt = pc.acc_scale, pc.fmt
    p.y_f32 = _p(out_f32)
    p.y_hi = _p(out_planes.hi if out_planes is not None else None)
    p.y_lo = _p(out_planes.lo if out_planes is not None else None)
    p.out_H, p.out_W = oH, oW
    if out_planes is not None:
        p.out_cstride, p.out_coffset = out_planes.cstride, out_planes.coffset + out_coffset
        if out_f32 is not None:
            assert out_f32.shape[-1] == out_planes.cstride, "f32 and planes outputs must share the channel stride"
    else:
        p.out_cstride, p.out_coffset = out_f32.shape[-1], out_coffset
    p.oh_mul, p.oh_off, p.ow_mul, p.ow_off = ohm, oho, owm, owo
    p.tile_n, p.stages = tile_n, stages
    p.deconv_phases = int(deconv_phases)
    p.stats_ws = _p(stats_ws)
    p.acc_chunk_kb = ACC_CHUNK_KB
    if SPLIT_K and not direct:
        # split-K workspace (few tiles, deep K): owned by the layer and kept for its lifetime, so captured CUDA graphs keep
        # pointing at live memory; zero-filled once (the arrival counters), the kernel leaves it reusable
        need = _lib.load().shineon_conv2d_splitk_workspace_bytes(C.byref(p))
        if need:
            cache = pc.__dict__.setdefault("_sk_ws", {})
            key = (N, H, W, Ho, Wo, x.cpad, tile_n, str(dev))
            ws = cache.get(key)
            if ws is None or ws.numel() < need:
                ws = cache[key] = torch.zeros(need, dtype=torch.uint8, device=dev)
            p.splitk_ws, p.splitk_ws_bytes = _p(ws), ws.numel()
    assert stats_ws is None or (not direct and stats_ws.dtype == torch.float64 and stats_ws.numel() == 2 * N * pc.Cout)
    fn = _lib.load().shineon_conv2d_direct_fwd if direct else _lib.load().shineon_conv2d_igemm_fwd
    prof = PROFILE
    if prof is not None and not direct:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(fn(C.byref(p), _stream()), "shineon_conv2d_direct_fwd" if direct else "shineon_conv2d_igemm_fwd")
    if prof is not None and not direct:
        e1.record()
        prof.append((2.0 * N * Ho * Wo * pc.Cout * pc.kh * pc.kw * pc.Cin * (4 if deconv_phases else 1), e0, e1,
                     (N, H, W, pc.Cin, x.cpad, pc.Cout, pc.kh, pc.stride)))
    return out_f32, out_planes


# ----------------------------------------------------------------------------- layout / norm / pointwise
def nchw_to_planes(x0, x1=None, act=None, act_param=0.0, prec=None, out=None):
    x0 = _req(x0, name="x0")
    N, C0, H, W = x0.shape
    C1 = 0
    if x1 is not None:
        x1 = _req(x1, name="x1")
        C1 = x1.shape[1]
    if out is None:
        out = Planes(N, H, W, C0 + C1, prec=prec, device=x0.device)
    check(_lib.load().shineon_nchw_to_planes(_p(x0), C0, _p(x1), C1, out._ptr(out.hi), out._ptr(out.lo), N, H, W,
                                             out.cstride, ACT[act], float(act_param), out.fmt, _stream()),
          "shineon_nchw_to_planes")
    return out


def planes_to_nchw(x):
    """Planes (or a channel window) -> f32 NCHW [N,C,H,W]."""
    y = torch.empty(x.N, x.C, x.H, x.W, dtype=torch.float32, device=x.hi.device)
    check(_lib.load().shineon_planes_to_nchw(x._ptr(x.hi), x._ptr(x.lo), x.cstride, _p(y), x.N, x.H, x.W, x.C, x.fmt,
                                             _stream()), "shineon_planes_to_nchw")
    return y


def nchw_im2col_planes(x0, x1, kh, kw, stride, pad, act=None, act_param=0.0, prec=None):
    """im2col'd first-layer input: Planes [N,Ho,Wo,pad64(kh*kw*C)] with k = (fy*kw+fx)*C + c."""
    x0 = _req(x0, name="x0")
    N, C0, H, W = x0.shape
    C1 = 0
    if x1 is not None:
        x1 = _req(x1, name="x1")
        C1 = x1.shape[1]
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    out = Planes(N, Ho, Wo, kh * kw * (C0 + C1), prec=prec, device=x0.device, zero_pad=False)  # kernel writes all of kpad
    check(_lib.load().shineon_nchw_im2col_planes(_p(x0), C0, _p(x1), C1, _p(out.hi), _p(out.lo), N, H, W, kh, kw, stride,
                                                 pad, Ho, Wo, out.cpad, ACT[act], float(act_param), out.fmt, _stream()),
          "shineon_nchw_im2col_planes")
    return out


class Im2colConv:
    """A Conv2d with tiny Cin as a 1x1 GEMM over nchw_im2col_planes (weights reordered tap-major to match)."""

    # True: the kernel's producer warps build the im2col tile in shared memory (nothing materialised in HBM).
    # Measured on B200 (profiles/r01_first_layers.md) the gather is latency-bound with only 8 producer warps per SM
    # and loses to the standalone full-occupancy im2col pass + TMA GEMM, so the materialised path stays the default.
    FUSED = False

    def __init__(self, weight, bias, stride, pad, prec=None):
        Cout, Cin, kh, kw = weight.shape
        self.kh, self.kw, self.stride, self.pad = kh, kw, stride, pad
        w2 = weight.detach().float().permute(0, 2, 3, 1).reshape(Cout, kh * kw * Cin, 1, 1).contiguous()
        self.pc = PackedConv(w2, bias, stride=1, pad=0, prec=prec)

    def prepare(self, x0, x1=None, act=None, act_param=0.0):
        """Materialised im2col planes (reference / fallback path; `conv` below never writes them to HBM)."""
        return nchw_im2col_planes(x0, x1, self.kh, self.kw, self.stride, self.pad, act, act_param,
                                  prec=(self.pc.fmt, self.pc.w_lo is not None))

    def conv(self, x0, x1=None, *, scale=None, shift=None, pre_act=None, post_act=None, act_param=0.0, want_f32=False,
             want_planes=False, out_f32=None, out_planes=None, fused=None):
        """The whole layer: im2col + GEMM.  fused=None -> Im2colConv.FUSED."""
        if not (self.FUSED if fused is None else fused):
            a = self.prepare(x0, x1)
            return conv2d(a, self.pc, scale=scale, shift=shift, pre_act=pre_act, post_act=post_act, act_param=act_param,
                          want_f32=want_f32, want_planes=want_planes, out_f32=out_f32, out_planes=out_planes)
        x0 = _req(x0, name="x0")
        N, C0, H, W = x0.shape
        C1 = 0
        if x1 is not None:
            x1 = _req(x1, name="x1")
            C1 = x1.shape[1]
        pc = self.pc
        Ho, Wo = (H + 2 * self.pad - self.kh) // self.stride + 1, (W + 2 * self.pad - self.kw) // self.stride + 1
        dev = x0.device
        prec = (pc.fmt, pc.w_lo is not None)
        if want_f32 and out_f32 is None:
            out_f32 = torch.empty(N, Ho, Wo, pc.Cout, dtype=torch.float32, device=dev)
        if want_planes and out_planes is None:
            out_planes = Planes(N, Ho, Wo, pc.Cout, prec=prec, device=dev)
        p = Conv2dParams()
        p.N, p.H, p.W, p.cin_pad = N, H, W, pc.cin_pad
        p.w_hi, p.w_lo = _p(pc.w_hi), _p(pc.w_lo)
        p.Cout, p.kh, p.kw, p.stride, p.pad_h, p.pad_w = pc.Cout, self.kh, self.kw, self.stride, self.pad, self.pad
        p.Ho, p.Wo = Ho, Wo
        p.bias, p.scale, p.shift = _p(pc.bias), _p(scale), _p(shift)
        p.pre_act, p.post_act, p.act_param = ACT[pre_act], ACT[post_act], float(act_param)
        p.acc_scale, p.plane_fmt = pc.acc_scale, pc.fmt
        p.y_f32 = _p(out_f32)
        p.y_hi = _p(out_planes.hi if out_planes is not None else None)
        p.y_lo = _p(out_planes.lo if out_planes is not None else None)
        if out_planes is not None:
            p.out_cstride, p.out_coffset = out_planes.cstride, out_planes.coffset
            assert out_f32 is None or out_f32.shape[-1] == out_planes.cstride
        else:
            p.out_cstride, p.out_coffset = out_f32.shape[-1], 0
        prof = PROFILE
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        check(_lib.load().shineon_conv2d_im2col_fwd(C.byref(p), _p(x0), C0, _p(x1), C1, _stream()),
              "shineon_conv2d_im2col_fwd")
        if prof is not None:
            e1.record()
            prof.append((2.0 * N * Ho * Wo * pc.Cout * self.kh * self.kw * (C0 + C1), e0, e1,
                         (N, H, W, C0 + C1, pc.cin_pad, pc.Cout, self.kh, self.stride)))
        return out_f32, out_planes


class S2dConv:
    """A 4x4 / stride 2 / pad 1 Conv2d with few input channels as a 2x2 / stride 1 conv over shifted space-to-depth planes
    (shineon_nchw_s2d_planes): every input element is written once (im2col writes it four times).  Same interface as
    Im2colConv."""

    def __init__(self, weight, bias, prec=None):
        Cout, Cin, kh, kw = weight.shape
        assert (kh, kw) == (4, 4)
        self.Cin = Cin
        # fy = 2*ay + py, fx = 2*ax + px  ->  [Cout, (py, px, c), ay, ax]
        w2 = weight.detach().float().reshape(Cout, Cin, 2, 2, 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(Cout, 4 * Cin, 2, 2)
        self.pc = PackedConv(w2.contiguous(), bias, stride=1, pad=0, prec=prec)

    @staticmethod
    def applicable(weight, stride, pad):
        Cout, Cin, kh, kw = weight.shape
        return (kh, kw, stride, pad) == (4, 4, 2, 1) and 8 <= Cin <= 32

    def prepare(self, x0, x1=None):
        x0 = _req(x0, name="x0")
        N, C0, H, W = x0.shape
        C1 = 0
        if x1 is not None:
            x1 = _req(x1, name="x1")
            C1 = x1.shape[1]
        assert C0 + C1 == self.Cin and H % 2 == 0 and W % 2 == 0
        out = Planes(N, H // 2 + 1, W // 2 + 1, 4 * self.Cin, prec=(self.pc.fmt, self.pc.w_lo is not None), device=x0.device,
                     zero_pad=False)  # the kernel writes the padding channels too
        check(_lib.load().shineon_nchw_s2d_planes(_p(x0), C0, _p(x1), C1, _p(out.hi), _p(out.lo), N, H, W, out.cpad, out.fmt,
                                                  _stream()), "shineon_nchw_s2d_planes")
        return out

    def conv(self, x0, x1=None, *, scale=None, shift=None, pre_act=None, post_act=None, act_param=0.0, want_f32=False,
             want_planes=False, out_f32=None, out_planes=None, fused=None):
        a = self.prepare(x0, x1)
        return self.conv_planes(a, scale=scale, shift=shift, pre_act=pre_act, post_act=post_act, act_param=act_param,
                                want_f32=want_f32, want_planes=want_planes, out_f32=out_f32, out_planes=out_planes)

    def conv_planes(self, a, **kw):
        """The GEMM on space-to-depth planes that already exist (prepare(), or written directly by the frame prep)."""
        return conv2d(a, self.pc, **kw)


def first_layer_conv(weight, bias, stride, pad, prec=None):
    """Small-Cin stem: space-to-depth form where it applies (4x4 s2 p1, 8 <= Cin <= 32), else im2col."""
    if S2dConv.applicable(weight, stride, pad):
        return S2dConv(weight, bias, prec=prec)
    return Im2colConv(weight, bias, stride, pad, prec=prec)


class TapStackedConv3x3:
    """3x3 / stride 1 / pad 1 conv with few output channels as one 1x1 GEMM with 9*Cout outputs + col2im."""

    def __init__(self, weight, bias, prec=None, cin_pad=None, chan_map=None):
        Cout, Cin, kh, kw = weight.shape
        assert kh == 3 and kw == 3
        self.Cout = Cout
        w2 = weight.detach().float().permute(2, 3, 0, 1).reshape(9 * Cout, Cin, 1, 1).contiguous()  # (tap, co) major
        self.pc = PackedConv(w2, None, stride=1, pad=0, prec=prec, cin_pad=cin_pad, chan_map=chan_map)
        self.bias = None if bias is None else _req(bias.detach().float().contiguous(), name="bias")

    def __call__(self, x):
        t, _ = conv2d(x, self.pc, want_f32=True)  # [N,H,W,9*Cout]
        y = torch.empty(x.N, x.H, x.W, self.Cout, dtype=torch.float32, device=t.device)
        check(_lib.load().shineon_col2im3x3(_p(t), _p(self.bias), _p(y), x.N, x.H, x.W, self.Cout, t.shape[-1], _stream()),
              "shineon_col2im3x3")
        return y


class UpsampledConv3x3:
    """nn.Upsample(x2, bilinear) -> Conv2d(3x3, pad 1) (unet.py:138-146) evaluated at the LOW resolution: the channel
    contraction commutes with the interpolation, so one tap-stacked 1x1 GEMM over the low-res planes (9*Cout outputs per
    pixel, a quarter of the FLOPs of the conv over the upsampled tensor, which is never materialised) is followed by
    `upconv3x3_gather`, which interpolates and sums the nine shifted partials with the exact border rules."""

    def __init__(self, weight, bias, prec=None, cin_pad=None, chan_map=None):
        Cout, Cin, kh, kw = weight.shape
        assert kh == 3 and kw == 3
        self.Cout, self.Cin = Cout, Cin
        w2 = weight.detach().float().permute(2, 3, 0, 1).reshape(9 * Cout, Cin, 1, 1).contiguous()  # (tap, co) major
        self.pc = PackedConv(w2, None, stride=1, pad=0, prec=prec, cin_pad=cin_pad, chan_map=chan_map)
        self.bias = None if bias is None else _req(bias.detach().float().contiguous(), name="bias")

    def __call__(self, x, stats_ws=None):
        """x: low-res Planes [N,h,w,cin_pad] -> f32 NHWC [N,2h,2w,Cout].  stats_ws: see conv2d."""
        global PROFILE
        prof, PROFILE = PROFILE, None  # one profile record for GEMM + gather, with the reference formulation's FLOPs
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        try:
            t, _ = conv2d(x, self.pc, want_f32=True)  # [N,h,w,9*Cout]
        finally:
            PROFILE = prof
        y = torch.empty(x.N, 2 * x.H, 2 * x.W, self.Cout, dtype=torch.float32, device=t.device)
        check(_lib.load().shineon_upconv3x3_gather(_p(t), _p(self.bias), _p(y), _p(stats_ws), x.N, x.H, x.W, self.Cout,
                                                   t.shape[-1], _stream()), "shineon_upconv3x3_gather")
        if prof is not None:
            e1.record()
            prof.append((2.0 * x.N * 4 * x.H * x.W * self.Cout * 9 * self.Cin, e0, e1,
                         (x.N, 2 * x.H, 2 * x.W, self.Cin, x.cpad, self.Cout, 3, 1),
                         2.0 * x.N * x.H * x.W * self.Cout * 9 * self.Cin))  # [4] = FLOPs actually issued
        return y


def instnorm_act(x, *, do_norm=True, act=None, act_param=0.0, eps=1e-5, want_f32=False, want_planes=True,
                 prec=None, out_f32=None, out_planes=None, ws=None, stats_ready=False):
    """x: f32 NHWC [N,H,W,C].  stats_ready: `ws` already holds the statistics (accumulated by x's producer)."""
    assert not stats_ready or ws is not None
    x = _req(x, name="x")
    N, H, W, Cc = x.shape
    if want_f32 and out_f32 is None:
        out_f32 = torch.empty_like(x)
    if want_planes and out_planes is None:
        out_planes = Planes(N, H, W, Cc, prec=prec, device=x.device)
    if ws is None and do_norm:
        ws = torch.empty(N * Cc * 2, dtype=torch.float64, device=x.device)
    # out_planes may be a channel window of a wider (concat) buffer: pointer offset + the buffer's channel stride
    check(_lib.load().shineon_instnorm_act(_p(x), _p(out_f32), out_planes._ptr(out_planes.hi) if out_planes else _p(None),
                                           out_planes._ptr(out_planes.lo) if out_planes else _p(None), _p(ws), N, H, W, Cc,
                                           out_planes.cstride if out_planes else Cc, float(eps), int(bool(do_norm)),
                                           int(bool(stats_ready)), ACT[act], float(act_param),
                                           out_planes.fmt if out_planes else 0, _stream()), "shineon_instnorm_act")
    return out_f32, out_planes


def upsample2x_cat(s0, s1=None, act=None, act_param=0.0, out=None):
    N, H, W = s0.N, s0.H, s0.W
    c1pad = s1.cpad if s1 is not None else 0
    if out is None:
        out = Planes(N, 2 * H, 2 * W, s0.cpad + c1pad, prec=s0.prec, device=s0.hi.device, cpad=s0.cpad + c1pad)
    check(_lib.load().shineon_upsample2x_cat(_p(s0.hi), _p(s0.lo), s0.cpad, _p(s1.hi if s1 else None),
                                             _p(s1.lo if s1 else None), c1pad, _p(out.hi), _p(out.lo), N, H, W,
                                             ACT[act], float(act_param), s0.fmt, _stream()), "shineon_upsample2x_cat")
    return out


def sagan_attention(qkv, x, gamma, Cq, *, act=None, act_param=0.0, want_f32=False, want_planes=True, prec=None,
                    out_f32=None, out_planes=None):
    """qkv: f32 NHWC [N,H,W,2*Cq+C]; x: f32 NHWC [N,H,W,C]."""
    qkv, x, gamma = _req(qkv), _req(x), _req(gamma)
    N, H, W, Cc = x.shape
    assert qkv.shape[-1] == 2 * Cq + Cc
    if want_f32 and out_f32 is None:
        out_f32 = torch.empty_like(x)
    if want_planes and out_planes is None:
        out_planes = Planes(N, H, W, Cc, prec=prec, device=x.device)
    check(_lib.load().shineon_sagan_attention(_p(qkv), _p(x), _p(gamma), _p(out_f32),
                                              out_planes._ptr(out_planes.hi) if out_planes else _p(None),
                                              out_planes._ptr(out_planes.lo) if out_planes else _p(None), N, H * W, Cc, Cq,
                                              out_planes.cstride if out_planes else Cc, ACT[act], float(act_param),
                                              out_planes.fmt if out_planes else 0, _stream()), "shineon_sagan_attention")
    return out_f32, out_planes


def feature_l2norm(x):
    """FeatureL2Norm.forward (warp.py:43-50): f32 NCHW -> x / sqrt(sum_c x^2 + 1e-6)."""
    x = _req(x, name="feature")
    B, Cc, H, W = x.shape
    y = torch.empty_like(x)
    check(_lib.load().shineon_feature_l2norm(_p(x), _p(y), B, Cc, H, W, _stream()), "shineon_feature_l2norm")
    return y


def l2norm_correlation(featA, featB, *, want_f32=False, want_planes=True, prec=None, normalize=True):
    """featA/B: f32 NHWC [B,h,w,C] -> correlation [B,h,w,h*w] (channel = wA*h+hA)."""
    featA, featB = _req(featA), _req(featB)
    B, h, w, Cc = featA.shape
    corr = torch.empty(B, h, w, h * w, dtype=torch.float32, device=featA.device) if want_f32 else None
    planes = Planes(B, h, w, h * w, prec=prec, device=featA.device) if want_planes else None
    check(_lib.load().shineon_l2norm_correlation(_p(featA), _p(featB), _p(corr), _p(planes.hi if planes else None),
                                                 _p(planes.lo if planes else None), B, h, w, Cc,
                                                 planes.cpad if planes else h * w, planes.fmt if planes else 0,
                                                 int(bool(normalize)), _stream()),
          "shineon_l2norm_correlation")
    return corr, planes


# tensor-core form of l2norm_correlation (per-image tcgen05 GEMM); False = the fp32 CUDA-core kernel
CORRELATION_TC = True


def l2norm_correlation_tc(featA, featB, prec=None, want_f32=False):
    """FeatureL2Norm x2 + FeatureCorrelation (warp.py:39-67) on the tensor cores: one pass writes both normalised feature
    maps as 16-bit planes (featA's rows in the transposed pixel order of warp.py:60), then the all-pairs products are ONE
    per-image implicit GEMM (featA's planes play the weights).  featA/B: f32 NHWC [B,h,w,C], C % 64 == 0.
    Returns (f32 NHWC [B,h,w,h*w] | None, Planes [B,h,w,h*w])."""
    featA, featB = _req(featA), _req(featB)
    B, h, w, Cc = featA.shape
    P = h * w
    fmt, split = resolve_precision(prec)
    dev = featA.device
    a = Planes(B, 1, P, Cc, prec=(fmt, split), device=dev, zero_pad=False)   # [B][P rows][C]
    b = Planes(B, h, w, Cc, prec=(fmt, split), device=dev, zero_pad=False)
    scale = 64.0
    check(_lib.load().shineon_l2norm_planes(_p(featA), _p(featB), _p(a.hi), _p(a.lo), _p(b.hi), _p(b.lo), B, h, w, Cc, fmt,
                                            scale, _stream()), "shineon_l2norm_planes")
    pc = PackedConv.__new__(PackedConv)
    pc.fmt = fmt
    pc.Cout, pc.Cin, pc.kh, pc.kw, pc.stride, pc.pad_h, pc.pad_w = P, Cc, 1, 1, 1, 0, 0
    pc.cin_pad, pc.w_hi, pc.w_lo, pc.bias, pc.acc_scale, pc.transposed, pc.per_image = Cc, a.hi, a.lo, None, 1.0 / (scale * scale), False, True
    return conv2d(b, pc, want_f32=want_f32, want_planes=True)


def linear_tanh(x, weight, bias):
    """x: f32 NHWC [B,h,w,C]; weight [out, C*h*w] over the NCHW-flattened input (warp.py:94-99)."""
    x, weight = _req(x), _req(weight)
    B, h, w, Cc = x.shape
    out_dim = weight.shape[0]
    assert weight.shape[1] == Cc * h * w
    theta = torch.empty(B, out_dim, dtype=torch.float32, device=x.device)
    check(_lib.load().shineon_linear_tanh(_p(x), _p(weight), _p(bias), _p(theta), B, h, w, Cc, out_dim, _stream()),
          "shineon_linear_tanh")
    return theta


def tom_compose(unet_out, cloth, n_frames, flow_warp, outs, frame=0, warped_prev=None, tryon_u8=None):
    """unet_out f32 NHWC [B,H,W,Cout]; outs = (p_rendereds, tryon_masks, p_tryons, flow_masks|None) NCHW, any may be None;
    tryon_u8: optional uint8 [B,n,H,W,3] receiving the frame's try-on image as visualization.save_images encodes it."""
    unet_out, cloth = _req(unet_out), _req(cloth)
    B, H, W, Cout = unet_out.shape
    pr, tm, pt, fm = outs
    if tryon_u8 is not None:
        _req(tryon_u8, torch.uint8, "tryon_u8")
        assert tuple(tryon_u8.shape) == (B, n_frames, H, W, 3)
    check(_lib.load().shineon_tom_compose(_p(unet_out), Cout, _p(cloth), _p(warped_prev), _p(pr), _p(tm), _p(pt),
                                          _p(fm), _p(tryon_u8), B, H, W, n_frames, frame, int(bool(flow_warp)), _stream()),
          "shineon_tom_compose")


def image_to_u8(x):
    """visualization.save_images' encoding (visualization.py:73-76) on the device: f32 [B,C,H,W] in [-1,1] ->
    uint8 [B,H,W,C] = clamp((x + 1) * 0.5 * 255, 0, 255) truncated.  Bit-exact on the same f32 input."""
    x = _req(x, name="image")
    B, Cc, H, W = x.shape
    y = torch.empty(B, H, W, Cc, dtype=torch.uint8, device=x.device)
    check(_lib.load().shineon_image_to_u8(_p(x), _p(y), B, Cc, H, W, _stream()), "shineon_image_to_u8")
    return y


# ----------------------------------------------------------------------------- SAMS generator passes
def chan_stats(x):
    """x: f32 NHWC [N,H,W,C] -> f64 [N*C*2] per-(image, channel) sum / sum of squares (InstanceNorm statistics)."""
    x = _req(x, name="x")
    N, H, W, Cc = x.shape
    ws = torch.empty(N * Cc * 2, dtype=torch.float64, device=x.device)
    check(_lib.load().shineon_chan_stats(_p(x), _p(ws), N, H * W, Cc, _stream()), "shineon_chan_stats")
    return ws


def spade_modulate(x, gb, *, stats_ws=None, nscale=None, nshift=None, eps=1e-5, act=None, act_param=0.0, want_f32=False,
                   want_planes=True, prec=None, out_f32=None, out_f32_coffset=0, out_planes=None):
    """SPADE.forward's `norm(x) * (1 + gamma) + beta` (+ the activation applied next), spade.py:68-84.
    x: f32 NHWC [N,H,W,C]; gb: f32 NHWC [N,H,W,>=2C] = (1 + gamma | beta) from one stacked conv.
    stats_ws -> InstanceNorm2d; nscale/nshift [C] -> eval-mode BatchNorm2d(affine=False); neither -> no normalisation.
    out_f32 may be a wider NHWC buffer (the stacked tensor of AttentiveMultiSpade): written at channel out_f32_coffset."""
    x, gb = _req(x, name="x"), _req(gb, name="gamma_beta")
    N, H, W, Cc = x.shape
    assert gb.shape[:3] == x.shape[:3] and gb.shape[3] >= 2 * Cc
    if want_f32 and out_f32 is None:
        out_f32 = torch.empty_like(x)
    if want_planes and out_planes is None:
        out_planes = Planes(N, H, W, Cc, prec=prec, device=x.device)
    mode = 1 if stats_ws is not None else (2 if nscale is not None else 0)
    yf = C_void(0) if out_f32 is None else C_void(_req(out_f32, name="out_f32").data_ptr() + 4 * out_f32_coffset)
    check(_lib.load().shineon_spade_modulate(_p(x), _p(stats_ws), _p(nscale), _p(nshift), _p(gb), gb.shape[3], yf,
                                             out_f32.shape[3] if out_f32 is not None else Cc,
                                             out_planes._ptr(out_planes.hi) if out_planes else _p(None),
                                             out_planes._ptr(out_planes.lo) if out_planes else _p(None), N, H, W, Cc,
                                             out_planes.cstride if out_planes else Cc, float(eps), mode, ACT[act],
                                             float(act_param), out_planes.fmt if out_planes else 0, _stream()),
          "shineon_spade_modulate")
    return out_f32, out_planes


def nearest_resize_nhwc(x, scale_factor):
    """nn.Upsample(scale_factor) (nearest) of an f32 NHWC tensor (sams_generator.py:295-310)."""
    x = _req(x, name="x")
    N, Hs, Ws, Cc = x.shape
    H, W = int(Hs * scale_factor), int(Ws * scale_factor)  # floor, like F.interpolate's output size
    y = torch.empty(N, H, W, Cc, dtype=torch.float32, device=x.device)
    check(_lib.load().shineon_nearest_resize_nhwc(_p(x), _p(y), N, Hs, Ws, H, W, Cc, 1.0 / scale_factor, 1.0 / scale_factor,
                                                  _stream()), "shineon_nearest_resize_nhwc")
    return y


def nearest_resize_planes(x, size, prec=None):
    """F.interpolate(x, size, mode="nearest") of an f32 NCHW tensor, as conv-operand Planes [N,H,W,pad64(C)] (spade.py:74)."""
    x = _req(x, name="segmap")
    N, Cc, Hs, Ws = x.shape
    H, W = size
    out = Planes(N, H, W, Cc, prec=prec, device=x.device, zero_pad=False)  # the kernel writes the padding channels
    check(_lib.load().shineon_nearest_resize_planes(_p(x), N, Cc, Hs, Ws, _p(out.hi), _p(out.lo), H, W, out.cpad, Hs / H, Ws / W,
                                                    out.fmt, _stream()), "shineon_nearest_resize_planes")
    return out


def nearest_im2col_planes(x, size, ks, prec=None):
    """im2col (ks x ks, stride 1, pad ks/2) of F.interpolate(x, size, mode="nearest"): Planes [N,H,W,pad64(ks*ks*C)]."""
    x = _req(x, name="segmap")
    N, Cc, Hs, Ws = x.shape
    H, W = size
    out = Planes(N, H, W, ks * ks * Cc, prec=prec, device=x.device, zero_pad=False)  # the kernel writes all of kpad
    check(_lib.load().shineon_nearest_im2col_planes(_p(x), N, Cc, Hs, Ws, _p(out.hi), _p(out.lo), H, W, ks, out.cpad, Hs / H,
                                                    Ws / W, out.fmt, _stream()), "shineon_nearest_im2col_planes")
    return out


def add_nhwc(a, b, out=None):
    a, b = _req(a, name="a"), _req(b, name="b")
    assert a.shape == b.shape
    if out is None:
        out = torch.empty_like(a)
    check(_lib.load().shineon_add_nhwc(_p(a), _p(b), _p(out), a.numel(), _stream()), "shineon_add_nhwc")
    return out


def sams_flow_blend(gen_out, out_frame, warped_prev=None):
    """gen_out: f32 NHWC [B,H,W,3|4]; out_frame: f32 [B,3,H,W] view of one frame slot (batch-strided); see the C ABI."""
    gen_out = _req(gen_out, name="gen_out")
    B, H, W, Cg = gen_out.shape
    assert out_frame.is_cuda and out_frame.dtype == torch.float32 and tuple(out_frame.shape) == (B, 3, H, W)
    assert out_frame.stride()[1:] == (H * W, W, 1), "frame slot must be dense per image"
    if warped_prev is not None:
        warped_prev = _req(warped_prev, name="warped_prev")
    check(_lib.load().shineon_sams_flow_blend(_p(gen_out), Cg, _p(warped_prev), _p(out_frame), out_frame.stride(0), B, H, W,
                                              _stream()), "shineon_sams_flow_blend")
    return out_frame


# ----------------------------------------------------------------------------- FlowNet2 glue
def flownet_normalize(inputs, rgb_max=1.0):
    inputs = _req(inputs, name="inputs")
    B, C3, F2, H, W = inputs.shape
    assert C3 == 3 and F2 == 2
    x = torch.empty(B, 6, H, W, dtype=torch.float32, device=inputs.device)
    ws = torch.empty(3 * B, dtype=torch.float64, device=inputs.device)
    check(_lib.load().shineon_flownet_normalize(_p(inputs), _p(x), _p(ws), B, H, W, float(rgb_max), _stream()),
          "shineon_flownet_normalize")
    return x


def upsample4x_flow(src_nhwc, mul, bilinear):
    src_nhwc = _req(src_nhwc, name="flow")
    B, h, w, cs = src_nhwc.shape
    dst = torch.empty(B, 2, 4 * h, 4 * w, dtype=torch.float32, device=src_nhwc.device)
    check(_lib.load().shineon_upsample4x_flow(_p(src_nhwc), cs, _p(dst), B, h, w, float(mul), int(bool(bilinear)),
                                              _stream()), "shineon_upsample4x_flow")
    return dst


def flow_deconv4x4s2_planes(flow_nhwc, weight, bias, out_planes):
    """ConvTranspose2d(2, 2, 4, 2, 1) of a predicted flow (f32 NHWC [B,h,w,>=2]) written into the first two channels of
    out_planes (a Planes / channel window [B,2h,2w,.]); weight f32 [2,2,4,4] CUDA (the module's layout), bias [2] or None."""
    flow_nhwc, weight = _req(flow_nhwc, name="flow"), _req(weight, name="weight")
    B, h, w, cs = flow_nhwc.shape
    assert tuple(weight.shape) == (2, 2, 4, 4) and out_planes.N == B and out_planes.H == 2 * h and out_planes.W == 2 * w
    check(_lib.load().shineon_flow_deconv4x4s2_planes(_p(flow_nhwc), cs, _p(weight), _p(bias), out_planes._ptr(out_planes.hi),
                                                      out_planes._ptr(out_planes.lo), out_planes.cstride, B, h, w, out_planes.fmt,
                                                      _stream()), "shineon_flow_deconv4x4s2_planes")
    return out_planes


def flownet_warp_concat(x, flow, div_flow):
    x, flow = _req(x), _req(flow)
    B, _, H, W = x.shape
    out = torch.empty(B, 12, H, W, dtype=torch.float32, device=x.device)
    check(_lib.load().shineon_flownet_warp_concat(_p(x), _p(flow), _p(out), B, H, W, float(div_flow), _stream()),
          "shineon_flownet_warp_concat")
    return out


def flownet_fusion_concat(x, flow_sd, flow_s2):
    x, flow_sd, flow_s2 = _req(x), _req(flow_sd), _req(flow_s2)
    B, _, H, W = x.shape
    out = torch.empty(B, 11, H, W, dtype=torch.float32, device=x.device)
    check(_lib.load().shineon_flownet_fusion_concat(_p(x), _p(flow_sd), _p(flow_s2), _p(out), B, H, W, _stream()),
          "shineon_flownet_fusion_concat")
    return out


def bilinear_resize(x, size, mul=1.0):
    """nn.Upsample(size=size, mode="bilinear") (align_corners=False) of an f32 NCHW tensor, times `mul`."""
    x = _req(x, name="x")
    B, Cc, Hi, Wi = x.shape
    Ho, Wo = size
    y = torch.empty(B, Cc, Ho, Wo, dtype=torch.float32, device=x.device)
    check(_lib.load().shineon_bilinear_resize(_p(x), _p(y), B * Cc, Hi, Wi, Ho, Wo, float(mul), _stream()),
          "shineon_bilinear_resize")
    return y


def flow_confidence(im1, im2, flow, threshold=0.02):
    im1, im2, flow = _req(im1), _req(im2), _req(flow)
    B, Cc, H, W = im1.shape
    conf = torch.empty(B, 1, H, W, dtype=torch.float32, device=im1.device)
    check(_lib.load().shineon_flow_confidence(_p(im1), _p(im2), _p(flow), _p(conf), B, Cc, H, W, float(threshold),
                                              _stream()), "shineon_flow_confidence")
    return conf


# ----------------------------------------------------------------------------- optimiser building block (rows U6/U7)
def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
              grad_scale=1.0):
    """In-place fused Adam on flat f32 CUDA buffers (torch.optim.Adam semantics)."""
    param, grad, exp_avg, exp_avg_sq = _req(param), _req(grad), _req(exp_avg), _req(exp_avg_sq)
    n = param.numel()
    assert grad.numel() == n and exp_avg.numel() == n and exp_avg_sq.numel() == n
    check(_lib.load().shineon_adam_step(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), n, float(lr), float(betas[0]),
                                        float(betas[1]), float(eps), float(weight_decay), int(step), float(grad_scale),
                                        _stream()), "shineon_adam_step")


# ----------------------------------------------------------------------------- conv backward (row U6)
_WGRAD_WS = {}


def _workspace(nbytes, device, key="ws"):
    """Grow-only scratch buffer per (device, key): the C ABI never allocates, the caller owns the workspace."""
    # one buffer per stream role (main / auxiliary): the weight-gradient branch of the backward runs on an auxiliary stream next to kernels that
    # use the same kind of scratch on the main one (networks/_engine_util.side_run)
    from .networks._engine_util import on_aux_stream

    k = (str(device), key, on_aux_stream(device))
    buf = _WGRAD_WS.get(k)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _WGRAD_WS[k] = buf
    return buf


def conv2d_wgrad(g, x, grad_w, *, Cout, Cin, kh, kw, stride, pad, mode=0, chan_map=None, alpha=1.0, beta=0.0,
                 out_hw=None, splits=0, desc_variant=0, g_coffset=0):
    """Weight gradient of a conv layer: g = Planes of dL/d(conv output) [N,Ho,Wo,>=Cout], x = the conv's input Planes.
    Writes beta*grad_w + alpha*dW into grad_w (f32, the parameter's own layout: OIHW for mode 0/1, IOHW for mode 2)."""
    from ._lib import Conv2dWgradParams

    assert g.fmt == x.fmt and (g.lo is None) == (x.lo is None), "precision of gradient and activation planes must agree"
    grad_w = _req(grad_w, name="grad_w")
    pad_h, pad_w = pad if isinstance(pad, tuple) else (pad, pad)
    p = Conv2dWgradParams()
    p.g_hi, p.g_lo, p.g_cstride = g._ptr(g.hi), g._ptr(g.lo), g.cstride
    p.g_coffset = g_coffset
    p.g_cpad = g.cpad if g_coffset == 0 else min(cpad64(Cout), (g.cstride - g_coffset) // 64 * 64)
    assert p.g_cpad >= 64, "G channel window too close to the end of the pixel row"
    p.x_hi, p.x_lo = x._ptr(x.hi), x._ptr(x.lo)
    p.N, p.H, p.W, p.cin_pad, p.x_cstride = x.N, x.H, x.W, x.cpad, x.cstride
    p.Cout, p.Cin, p.kh, p.kw, p.stride, p.pad_h, p.pad_w = Cout, Cin, kh, kw, stride, pad_h, pad_w
    p.Ho, p.Wo = out_hw if out_hw is not None else (g.H, g.W)
    assert g.N == x.N and (g.H, g.W) == (p.Ho, p.Wo)
    p.plane_fmt = g.fmt
    cm = None
    if chan_map is not None:
        cm = chan_map if isinstance(chan_map, torch.Tensor) else torch.as_tensor(chan_map, dtype=torch.int32, device=grad_w.device)
        assert cm.numel() == x.cpad and cm.dtype == torch.int32
    p.chan_map = _p(cm)
    p.mode, p.grad_w, p.alpha, p.beta = mode, _p(grad_w), float(alpha), float(beta)
    p.splits, p.desc_variant = splits, desc_variant
    lib = _lib.load()
    need = lib.shineon_conv2d_wgrad_workspace_bytes(C.byref(p))
    ws = _workspace(need, grad_w.device)
    p.workspace, p.workspace_bytes = _p(ws), ws.numel()
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib.shineon_conv2d_wgrad(C.byref(p), _stream()), "shineon_conv2d_wgrad")
    if prof is not None:
        e1.record()
        prof.append((2.0 * x.N * p.Ho * p.Wo * Cout * kh * kw * Cin, e0, e1, ("wgrad", x.N, x.H, x.W, Cin, x.cpad, Cout, kh, stride)))
    return grad_w


def channel_sum(x, grad, alpha=1.0, beta=0.0, coffset=0):
    """grad[c] = beta*grad[c] + alpha * sum over all pixels of the NHWC f32 tensor x[..., coffset + c]  (bias gradient)."""
    x, grad = _req(x, name="x"), _req(grad, name="grad")
    Cc = grad.numel()
    cs = x.shape[-1]
    assert cs >= coffset + Cc
    ws = _workspace(8 * Cc, x.device, key="chsum")
    check(_lib.load().shineon_channel_sum(C_void(x.data_ptr() + 4 * coffset), _p(grad), _p(ws), x.numel() // cs, Cc, cs, float(alpha), float(beta),
                                          _stream()), "shineon_channel_sum")
    return grad


def instnorm_stats_ws(x):
    """Statistics workspace of instnorm_act for an NHWC tensor (kept by the training path for the backward)."""
    N, _, _, Cc = x.shape
    return torch.empty(N * Cc * 2, dtype=torch.float64, device=x.device)


def instnorm_act_bwd(x, stats_fwd, g1, g2=None, *, do_norm=True, act=None, act_param=0.0, eps=1e-5, want_f32=False,
                     want_planes=True, prec=None):
    """Backward of instnorm_act.  x: the f32 NHWC conv output the forward normalised; g1 (+g2): dL/d(activated output).
    Returns (gx_f32|None, gx Planes|None)."""
    x, g1 = _req(x, name="x"), _req(g1, name="g1")
    if g2 is not None:
        g2 = _req(g2, name="g2")
        assert g2.shape == x.shape
    assert g1.shape == x.shape
    N, H, W, Cc = x.shape
    gx = torch.empty_like(x) if want_f32 else None
    gp = Planes(N, H, W, Cc, prec=prec, device=x.device) if want_planes else None
    ws = torch.empty(N * Cc * 2, dtype=torch.float64, device=x.device) if do_norm else None
    check(_lib.load().shineon_instnorm_act_bwd(_p(x), _p(stats_fwd), _p(g1), _p(g2), _p(gx), _p(gp.hi if gp else None),
                                               _p(gp.lo if gp else None), _p(ws), N, H, W, Cc, gp.cpad if gp else Cc,
                                               float(eps), int(bool(do_norm)), ACT[act], float(act_param),
                                               gp.fmt if gp else 0, _stream()), "shineon_instnorm_act_bwd")
    return gx, gp


def act_bwd(z, g1, g2=None, act=None, act_param=0.0):
    z, g1 = _req(z, name="z"), _req(g1, name="g1")
    if g2 is not None:
        g2 = _req(g2, name="g2")
    gz = torch.empty_like(z)
    check(_lib.load().shineon_act_bwd(_p(z), _p(g1), _p(g2), _p(gz), z.numel(), ACT[act], float(act_param), _stream()),
          "shineon_act_bwd")
    return gz


def upsample2x_cat_bwd(g_up, C0, C1=0):
    """g_up: f32 NHWC [N,2H,2W,>=C0+C1] -> (g0 [N,H,W,C0], g1 [N,H,W,C1] | None)."""
    g_up = _req(g_up, name="g_up")
    N, H2, W2, cs = g_up.shape
    H, W = H2 // 2, W2 // 2
    g0 = torch.empty(N, H, W, C0, dtype=torch.float32, device=g_up.device)
    g1 = torch.empty(N, H, W, C1, dtype=torch.float32, device=g_up.device) if C1 else None
    check(_lib.load().shineon_upsample2x_cat_bwd(_p(g_up), cs, _p(g0), C0, _p(g1), C1, N, H, W, _stream()),
          "shineon_upsample2x_cat_bwd")
    return g0, g1


def sagan_attention_bwd(qkv, gamma, g_out, Cq, g_gamma, beta_gamma=1.0):
    """-> g_qkv f32 [N,H,W,2Cq+C]; accumulates dL/dgamma into g_gamma (f32 [1])."""
    qkv, gamma, g_out, g_gamma = _req(qkv), _req(gamma), _req(g_out), _req(g_gamma)
    N, H, W, Cc = g_out.shape
    assert qkv.shape[-1] == 2 * Cq + Cc
    g_qkv = torch.empty_like(qkv)
    lib = _lib.load()
    need = lib.shineon_sagan_attention_bwd_workspace_bytes(N, H * W)
    ws = _workspace(need, qkv.device, key="attn_bwd")
    check(lib.shineon_sagan_attention_bwd(_p(qkv), _p(gamma), _p(g_out), _p(g_qkv), _p(g_gamma), _p(ws), ws.numel(), N,
                                          H * W, Cc, Cq, float(beta_gamma), _stream()), "shineon_sagan_attention_bwd")
    return g_qkv


def tom_compose_bwd(unet_out, cloth, n_frames, flow_warp, g_unet_out, frame=0, warped_prev=None, g_rendereds=None,
                    g_masks=None, g_tryons=None, g_flow_masks=None, want_g_warped=False):
    unet_out, cloth, g_unet_out = _req(unet_out), _req(cloth), _req(g_unet_out)
    B, H, W, Cout = unet_out.shape
    g_warped = torch.empty(B, 3, H, W, dtype=torch.float32, device=unet_out.device) if want_g_warped else None
    check(_lib.load().shineon_tom_compose_bwd(_p(unet_out), Cout, _p(cloth), _p(warped_prev), _p(g_rendereds), _p(g_masks),
                                              _p(g_tryons), _p(g_flow_masks), _p(g_unet_out), _p(g_warped), B, H, W,
                                              n_frames, frame, int(bool(flow_warp)), _stream()), "shineon_tom_compose_bwd")
    return g_warped


def l1_loss(a, b, loss, grad_a=None, weight=1.0, beta_loss=1.0, accumulate_grad=False):
    """loss[0] = beta_loss*loss[0] + weight*mean|a-b|; grad_a (+)= weight*sign(a-b)/numel."""
    a, b, loss = _req(a, name="a"), _req(b, name="b"), _req(loss, name="loss")
    assert a.shape == b.shape
    ws = _workspace(8, a.device, key="l1")
    check(_lib.load().shineon_l1_loss(_p(a), _p(b), _p(grad_a), _p(loss), _p(ws), a.numel(), float(weight),
                                      float(beta_loss), int(bool(accumulate_grad)), _stream()), "shineon_l1_loss")
    return loss


def maxpool2x2(x, want_f32=True, want_planes=True, prec=None):
    x = _req(x, name="x")
    N, H, W, Cc = x.shape
    y = torch.empty(N, H // 2, W // 2, Cc, dtype=torch.float32, device=x.device) if want_f32 else None
    yp = Planes(N, H // 2, W // 2, Cc, prec=prec, device=x.device) if want_planes else None
    check(_lib.load().shineon_maxpool2x2_fwd(_p(x), _p(y), _p(yp.hi if yp else None), _p(yp.lo if yp else None), N, H, W,
                                             Cc, yp.cpad if yp else Cc, yp.fmt if yp else 0, _stream()),
          "shineon_maxpool2x2_fwd")
    return y, yp


def maxpool2x2_bwd(x, g_y):
    x, g_y = _req(x, name="x"), _req(g_y, name="g_y")
    N, H, W, Cc = x.shape
    g_x = torch.empty_like(x)
    check(_lib.load().shineon_maxpool2x2_bwd(_p(x), _p(g_y), g_y.shape[-1], _p(g_x), N, H, W, Cc, _stream()),
          "shineon_maxpool2x2_bwd")
    return g_x


def nhwc_to_nchw_add(x, y, accumulate=True):
    """y [N,C,H,W] (+)= x [N,H,W,>=C]."""
    x, y = _req(x, name="x"), _req(y, name="y")
    N, Cc, H, W = y.shape
    assert x.shape[:3] == (N, H, W) and x.shape[3] >= Cc
    check(_lib.load().shineon_nhwc_to_nchw_add(_p(x), x.shape[3], _p(y), N, H, W, Cc, int(bool(accumulate)), _stream()),
          "shineon_nhwc_to_nchw_add")
    return y


# ----------------------------------------------------------------------------- dataset-side frame prep
class FramePrep:
    """TryonDataset.get_person_representation + get_cloth_representation (datasets/tryon_dataset.py:156-175,203-251,
    323-447) on the device, from decoded 8-bit frames (channel-last uint8 CUDA tensors).  Bit-identical to the
    reference's CPU tensors, including its two quirks: the cloth mask compares the normalised cloth with the 0-255
    threshold, and the 18 cocopose channels are the constant -1 (the squares only reach `im_cocopose`)."""

    def __init__(self, H=256, W=192, cloth_mask_threshold=240, radius=5, n_joints=18, device="cuda"):
        self.H, self.W, self.thr, self.radius, self.J = H, W, float(cloth_mask_threshold), int(radius), int(n_joints)
        lib = _lib.load()
        self._tabs = []
        for a, b in ((W, W // 16), (H, H // 16), (W // 16, W), (H // 16, H)):  # the two Image.resize calls, per axis
            ks = lib.shineon_pil_bilinear_coeffs(a, b, None, None)
            if ks <= 0:
                check(ks, "shineon_pil_bilinear_coeffs")
            bounds = torch.zeros(b, 2, dtype=torch.int32)
            kk = torch.zeros(b, ks, dtype=torch.int32)
            rc = lib.shineon_pil_bilinear_coeffs(a, b, C.c_void_p(bounds.data_ptr()), C.c_void_p(kk.data_ptr()))
            if rc != ks:
                check(rc if rc < 0 else -1, "shineon_pil_bilinear_coeffs")
            self._tabs.append((bounds.to(device), kk.to(device), ks))

    def __call__(self, parse, cloth, densepose, image, pose=None, want_image=False):
        """parse [F,H,W], cloth / densepose / image [F,H,W,3] uint8 CUDA; pose [F,J,3] float64 CUDA or None.
        Returns a dict with the batch keys the try-on stages read (f32 NCHW)."""
        for name, t, nd in (("parse", parse, 3), ("cloth", cloth, 4), ("densepose", densepose, 4), ("image", image, 4)):
            _req(t, torch.uint8, name)
            assert t.dim() == nd and tuple(t.shape[1:3]) == (self.H, self.W), f"{name}: shape {tuple(t.shape)}"
        F = parse.shape[0]
        dev = parse.device
        f32 = lambda c: torch.empty(F, c, self.H, self.W, dtype=torch.float32, device=dev)
        out = {"cloth": f32(3), "cloth_mask": f32(1), "densepose": f32(3), "agnostic": f32(4), "cocopose": f32(self.J)}
        if want_image:
            out["image"] = f32(3)
        p = _lib.FramePrepParams()
        p.image, p.parse, p.cloth, p.densepose = _p(image), _p(parse), _p(cloth), _p(densepose)
        p.pose = _p(None)
        if pose is not None:
            pose = _req(pose, torch.float64, "pose")
            assert tuple(pose.shape) == (F, self.J, 3)
            p.pose = _p(pose)
            out["im_cocopose"] = f32(1)
        p.image_out = _p(out.get("image"))
        p.cloth_out, p.cloth_mask_out, p.densepose_out = _p(out["cloth"]), _p(out["cloth_mask"]), _p(out["densepose"])
        p.agnostic_out, p.cocopose_out, p.im_cocopose_out = _p(out["agnostic"]), _p(out["cocopose"]), _p(out.get("im_cocopose"))
        for i, (b, k, ks) in enumerate(self._tabs):
            p.tab_bounds[i], p.tab_kk[i], p.tab_ksize[i] = b.data_ptr(), k.data_ptr(), ks
        p.F, p.H, p.W, p.n_joints, p.radius, p.cloth_mask_threshold = F, self.H, self.W, self.J, self.radius, self.thr
        check(_lib.load().shineon_frame_prep(C.byref(p), _stream()), "shineon_frame_prep")
        return out


class S2dInput:
    """Shifted space-to-depth planes of an f32 NCHW tensor [N,C,H,W] that never existed as such (FramePrep.planes +
    tps_warp_u8_planes write them directly): what S2dConv.prepare would have produced."""

    def __init__(self, planes, C, H, W):
        self.planes, self.C, self.H, self.W = planes, C, H, W
        self.N = planes.N


def frame_prep_planes(prep, parse, cloth, densepose, image, prec=None):
    """FramePrep fused with the first layers' layout passes (TryOnPipeline.run_raw): decoded 8-bit frames -> the three
    stem operands as 16-bit planes, bit-identical to prep(...) -> torch.cat -> S2dConv.prepare / Im2colConv.prepare:
      gmm_person  S2dInput of cat(agnostic, cocopose)              [F, 4+J, H, W]
      unet_in     S2dInput of cat(agnostic, densepose, cloth')     [F, 10, H, W]   (cloth' slots zero: tps_warp_u8_planes)
      cloth_i2c   Planes [F, H/2, W/2, 64]: im2col (4x4 s2 p1) of the cloth."""
    for name, t, nd in (("parse", parse, 3), ("cloth", cloth, 4), ("densepose", densepose, 4), ("image", image, 4)):
        _req(t, torch.uint8, name)
        assert t.dim() == nd and tuple(t.shape[1:3]) == (prep.H, prep.W), f"{name}: shape {tuple(t.shape)}"
    F, H, W, J = parse.shape[0], prep.H, prep.W, prep.J
    dev = parse.device
    fmt, split = resolve_precision(prec)
    pg = Planes(F, H // 2 + 1, W // 2 + 1, 4 * (4 + J), prec=(fmt, split), device=dev, zero_pad=False)
    pu = Planes(F, H // 2 + 1, W // 2 + 1, 40, prec=(fmt, split), device=dev, zero_pad=False)
    pc = Planes(F, H // 2, W // 2, 48, prec=(fmt, split), device=dev, zero_pad=False)
    sil = torch.empty(F, H, W, dtype=torch.uint8, device=dev)
    p = _lib.FramePrepPlanesParams()
    p.image, p.parse, p.cloth, p.densepose, p.silhouette_scratch = _p(image), _p(parse), _p(cloth), _p(densepose), _p(sil)
    p.gmm_hi, p.gmm_lo, p.unet_hi, p.unet_lo, p.cloth_hi, p.cloth_lo = _p(pg.hi), _p(pg.lo), _p(pu.hi), _p(pu.lo), _p(pc.hi), _p(pc.lo)
    p.gmm_cpad, p.unet_cpad, p.cloth_cpad, p.plane_fmt = pg.cpad, pu.cpad, pc.cpad, fmt
    for i, (b, k, ks) in enumerate(prep._tabs):
        p.tab_bounds[i], p.tab_kk[i], p.tab_ksize[i] = b.data_ptr(), k.data_ptr(), ks
    p.F, p.H, p.W, p.n_joints = F, H, W, J
    check(_lib.load().shineon_frame_prep_planes(C.byref(p), _stream()), "shineon_frame_prep_planes")
    return {"gmm_person": S2dInput(pg, 4 + J, H, W), "unet_in": S2dInput(pu, 10, H, W), "cloth_i2c": pc}


def tps_warp_u8_planes(theta, tables, cloth_u8, unet_in, c_off=7):
    """TPS warp (border padding) of the decoded 8-bit cloth [B,H,W,3]: returns the f32 NCHW warped cloth and fills the cloth'
    channels [c_off, c_off+3) of the U-Net stem operand `unet_in` (S2dInput)."""
    theta = _req(theta, name="theta")
    _req(cloth_u8, torch.uint8, "cloth_u8")
    B, H, W, _ = cloth_u8.shape
    z = unet_in.planes
    assert (unet_in.H, unet_in.W, z.N) == (H, W, B) and z.coffset == 0
    warped = torch.empty(B, 3, H, W, dtype=torch.float32, device=theta.device)
    check(_lib.load().shineon_tps_warp_u8_planes(_p(theta), C.byref(tables.struct), _p(cloth_u8), _p(warped), _p(z.hi), _p(z.lo),
                                                 z.cstride, unet_in.C, c_off, z.fmt, B, H, W, _stream()),
          "shineon_tps_warp_u8_planes")
    return warped


def flo_decode(flo_bytes, device="cuda"):
    """Middlebury .flo file content (bytes / uint8 tensor on the host) -> normalised flow f32 [2,H,W] on the device
    (flow_utils.readFlow + flow_norm, datasets/tryon_dataset.py:121,288-289).  Header errors raise like the reference."""
    import struct

    buf = torch.frombuffer(bytearray(flo_bytes), dtype=torch.uint8) if not isinstance(flo_bytes, torch.Tensor) else flo_bytes
    if buf.numel() < 12:
        raise ValueError("Invalid .flo file: truncated header")
    magic, w, h = struct.unpack("<fii", bytes(buf[:12].tolist()))
    if magic != 202021.25:
        raise ValueError("Magic number incorrect. Invalid .flo file")
    if buf.numel() < 12 + 8 * w * h:
        raise ValueError("Invalid .flo file: truncated payload")
    d = buf.to(device)
    out = torch.empty(2, h, w, dtype=torch.float32, device=device)
    check(_lib.load().shineon_flo_decode(C.c_void_p(d.data_ptr() + 12), _p(out), h, w, _stream()), "shineon_flo_decode")
    return out
