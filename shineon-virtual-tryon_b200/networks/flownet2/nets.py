"""FlowNet2 and its sub-networks (reference: models/flownet2_pytorch/models.py:32-192,
networks/FlowNetC.py, FlowNetS.py, FlowNetSD.py, FlowNetFusion.py, submodules.py) — same attribute names and
state_dict keys (`flownetc.conv1.0.weight`, `flownets_1.deconv5.0.weight`, `flownetfusion.predict_flow0.bias` ...).

Engine (eval mode, batchNorm=False as FlowNet2 is built by the reference, models.py:34): every Conv2d /
ConvTranspose2d runs on the tcgen05 implicit-GEMM kernel with the LeakyReLU(0.1) in its epilogue; `torch.cat`s
never happen — producers write straight into 64-aligned channel windows of one concat buffer and the consumer's
packed weights follow that layout; the small-Cin stems (3/6/11/12 channels, up to 7x7) use the im2col'd 1x1 GEMM;
the warps / norms / concats between sub-networks are the fused kernels of csrc/flownet_glue.cu.
"""
import os

import torch
import torch.nn as nn

from ... import ops
from .._engine_util import params_signature, require_cuda
from ..deconv import PackedDeconv4x4s2
from .native_ops import ChannelNorm, Correlation, Resample2d

LEAK = 0.1


# ------------------------------------------------------------------ module constructors (submodules.py:7-38)
def conv(batchNorm, in_planes, out_planes, kernel_size=3, stride=1):
    if batchNorm:
        raise NotImplementedError("FlowNet2 is built with batchNorm=False by the reference (models.py:34)")
    return nn.Sequential(
        nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=(kernel_size - 1) // 2, bias=True),
        nn.LeakyReLU(0.1, inplace=True))


def i_conv(batchNorm, in_planes, out_planes, kernel_size=3, stride=1, bias=True):
    if batchNorm:
        raise NotImplementedError("batchNorm=True")
    return nn.Sequential(nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride,
                                   padding=(kernel_size - 1) // 2, bias=bias))


def predict_flow(in_planes):
    return nn.Conv2d(in_planes, 2, kernel_size=3, stride=1, padding=1, bias=True)


def deconv(in_planes, out_planes):
    return nn.Sequential(nn.ConvTranspose2d(in_planes, out_planes, kernel_size=4, stride=2, padding=1, bias=True),
                         nn.LeakyReLU(0.1, inplace=True))


def _init(net):
    for m in net.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            if m.bias is not None:
                nn.init.uniform_(m.bias)
            nn.init.xavier_uniform_(m.weight)


# ------------------------------------------------------------------ engine helpers
class _Segs(list):
    """Segment channel counts of a concat buffer + the alignment of their offsets (64 = every segment readable on its own
    as a conv input window; 8 = packed for write-only windows)."""

    def __init__(self, seg_channels, align=64):
        super().__init__(seg_channels)
        self.align = align


def _seg_pad(c, align):
    return (c + align - 1) // align * align


class _Concat:
    """A concat buffer: segments at aligned channel offsets of one Planes; `chan_map` for the consumers.  align = 8 packs
    the segments (only the first one is then read on its own): FlowNetFusion's full-resolution concat of 64 + 16 + 2 channels
    is 128 wide instead of 192, its half-resolution one (128 + 32 + 2) 192 instead of 256 -- a third less operand traffic
    and K for the memory-bound layers that read them."""

    def __init__(self, N, H, W, seg_channels, prec, device, align=64):
        self.offsets, off = [], 0
        for c in seg_channels:
            self.offsets.append(off)
            off += _seg_pad(c, align)
        off = ops.cpad64(off)
        self.align = align
        self.seg = _Segs(seg_channels, align)
        self.buf = ops.Planes(N, H, W, off, prec=prec, device=device, cpad=off)
        self.buf.hi.zero_()
        if self.buf.lo is not None:
            self.buf.lo.zero_()

    def window(self, i):
        return self.buf.window(self.offsets[i], self.seg[i], align=self.align)

    @staticmethod
    def chan_map(seg_channels):
        align = getattr(seg_channels, "align", 64)
        cmap, base = [], 0
        for c in seg_channels:
            cmap += list(range(base, base + c)) + [-1] * (_seg_pad(c, align) - c)
            base += c
        return cmap + [-1] * (ops.cpad64(len(cmap)) - len(cmap))


# upsampled_flow* (2 -> 2 channel transposed convs on a predicted flow) on the direct CUDA-core kernel (A/B switch)
UPFLOW_DIRECT = os.environ.get("SHINEON_UPFLOW_DIRECT", "1") != "0"

# FlowNetFusion's concat buffers with packed (8-aligned) segments (A/B switch; 64 = one K-block per segment)
FUSION_CONCAT_ALIGN = 8 if os.environ.get("SHINEON_FUSION_TIGHT_CONCAT", "1") != "0" else 64

# forwards that may run concurrently (models.flownet.FlowNet compute lanes) must not share concat buffers: the lane id
# of the forward being issued is part of the cache key
CONCAT_LANE = [0]


class _Net(nn.Module):
    """Packed-weight cache shared by the four sub-networks."""

    def _concat(self, N, H, W, seg_channels, prec, device, align=64):
        """Concat buffers are kept per (shape, precision): every forward's producers overwrite all real channels of
        their windows and nothing ever writes the padding channels, so the zero fill happens once per buffer instead
        of once per forward (it was 78 fill launches = 5 % of a batch-16 FlowNet2 forward)."""
        cache = self.__dict__.setdefault("_cat_cache", {})
        key = (N, H, W, tuple(seg_channels), prec, str(device), CONCAT_LANE[0], align)
        cat = cache.get(key)
        if cat is None:
            if len(cache) >= 32:  # a new batch size / resolution: drop the old set
                cache.clear()
            cat = cache[key] = _Concat(N, H, W, seg_channels, prec, device, align)
        return cat

    # deconv{lvl} of a refinement level does not depend on that level's predict_flow -> upsampled_flow chain: it runs on an
    # auxiliary stream (a parallel branch under graph capture).  Both chains write windows of the cached concat buffer, so
    # no allocation crosses streams.
    parallel = os.environ.get("SHINEON_FLOW_PARALLEL_REFINE", "1") != "0"

    def _fork(self, fn, allocates=False):
        """Runs fn() on this network's auxiliary stream, forked from the current one; returns the join callable.
        allocates=True: fn's results are torch allocations that cross back to the forking stream -- only done inside a graph
        capture (whose private pool is not recycled before the join); eagerly such branches stay on the current stream."""
        if not self.parallel or ops.PROFILE is not None or (allocates and not torch.cuda.is_current_stream_capturing()):
            fn()
            return lambda: None
        cur = torch.cuda.current_stream()
        aux = self.__dict__.get("_aux_stream")
        if aux is None:
            aux = self.__dict__["_aux_stream"] = torch.cuda.Stream()
        aux.wait_stream(cur)
        with torch.cuda.stream(aux):
            fn()
        return lambda: cur.wait_stream(aux)

    def _packs(self, prec):
        sig = (params_signature(self), prec)
        if getattr(self, "_pk", None) is None or self._pk[0] != sig:
            require_cuda(self, type(self).__name__)
            self._pk = (sig, {})
        return self._pk[1]

    def _pc(self, prec, name, segs=None, stem=False):
        """PackedConv / Im2colConv / PackedDeconv4x4s2 for the attribute `name`."""
        pk = self._packs(prec)
        if name not in pk:
            m = getattr(self, name)
            m = m[0] if isinstance(m, nn.Sequential) else m
            cmap = _Concat.chan_map(segs) if segs is not None else None
            cin_pad = len(cmap) if cmap is not None else None
            if isinstance(m, nn.ConvTranspose2d):
                pk[name] = PackedDeconv4x4s2(m.weight, m.bias, cin_pad=cin_pad, prec=prec, chan_map=cmap)
            elif stem:
                pk[name] = ops.Im2colConv(m.weight, m.bias, m.stride[0], m.padding[0], prec=prec)
            else:
                pk[name] = ops.PackedConv(m.weight, m.bias, stride=m.stride[0], pad=m.padding[0], cin_pad=cin_pad,
                                          prec=prec, chan_map=cmap)
        return pk[name]

    def _tap(self, prec, name, segs=None):
        """3x3 / stride 1 / pad 1 layers with <= 16 output channels (every predict_flow, FlowNetFusion's inter convs) as a
        tap-stacked 1x1 GEMM + col2im: the activation is read once instead of once per tap from L2, and the K loop of
        the small-resolution predict_flow layers (2 CTAs, 144 sequential K-blocks: latency-bound) is 9x shorter."""
        m = getattr(self, name)
        m = m[0] if isinstance(m, nn.Sequential) else m
        if not (isinstance(m, nn.Conv2d) and m.kernel_size == (3, 3) and m.stride == (1, 1) and m.padding == (1, 1)
                and m.out_channels <= 16):
            return None
        pk = self._packs(prec)
        key = name + "/tap"
        if key not in pk:
            cmap = _Concat.chan_map(segs) if segs is not None else None
            pk[key] = ops.TapStackedConv3x3(m.weight, m.bias, prec=prec, cin_pad=len(cmap) if cmap is not None else None,
                                            chan_map=cmap)
        return pk[key]

    # conv + LeakyReLU(0.1) -> planes (optionally into a concat window)
    def _c(self, prec, name, x, out=None, segs=None, act=True):
        tap = self._tap(prec, name, segs)
        if tap is not None:
            _, y = ops.instnorm_act(tap(x), do_norm=False, act="leaky" if act else None, act_param=LEAK, want_f32=False,
                                    want_planes=True, prec=prec, out_planes=out)
            return y
        pc = self._pc(prec, name, segs)
        _, y = ops.conv2d(x, pc, post_act="leaky" if act else None, act_param=LEAK, want_planes=out is None,
                          out_planes=out)
        return y

    def _stem(self, prec, name, x_nchw, out=None):
        i2c = self._pc(prec, name, stem=True)
        _, y = i2c.conv(x_nchw, post_act="leaky", act_param=LEAK, want_planes=out is None, out_planes=out)
        return y

    def _flow(self, prec, name, x, segs=None, want_f32=False, want_planes=True):
        """predict_flow: 3x3 conv -> 2 channels, no activation; returns (f32 NHWC [B,h,w,2] | None, planes | None)."""
        tap = self._tap(prec, name, segs)
        if tap is not None:
            f32 = tap(x)
            pl = ops.instnorm_act(f32, do_norm=False, want_f32=False, want_planes=True, prec=prec)[1] if want_planes else None
            return (f32 if want_f32 else None), pl
        pc = self._pc(prec, name, segs)
        f32 = ops.conv2d(x, pc, want_f32=True)[0] if want_f32 else None
        pl = ops.conv2d(x, pc, want_planes=True)[1] if want_planes else None
        return f32, pl

    def _upflow(self, prec, name, flow_f32, flow_planes, out):
        """upsampled_flow*: ConvTranspose2d(2, 2, 4, 2, 1) of a predicted flow into `out` (the flow window of the next level's
        concat buffer).  From the f32 flow on the small CUDA-core kernel (16 MACs per output value); UPFLOW_DIRECT = False keeps
        the tensor-core transposed conv over the flow's 64-channel padded planes."""
        if UPFLOW_DIRECT and flow_f32 is not None:
            pk = self._packs(prec)
            key = name + "/f32"
            if key not in pk:
                m = getattr(self, name)
                pk[key] = (m.weight.detach().float().contiguous(), None if m.bias is None else m.bias.detach().float().contiguous())
            wgt, bias = pk[key]
            return ops.flow_deconv4x4s2_planes(flow_f32, wgt, bias, out)
        return self._pc(prec, name)(flow_planes, out_planes=out)

    def _refine(self, prec, c6, cats, names):
        """Decoder shared by FlowNetC / FlowNetS (FlowNetC.py:100-123): cats = [concat5, concat4, concat3, concat2]."""
        src, segs = c6, None
        flow = None
        for lvl, cat in zip((5, 4, 3, 2), cats):
            join = self._fork(lambda: self._pc(prec, f"deconv{lvl}", segs)(src, post_act="leaky", act_param=LEAK,
                                                                          out_planes=cat.window(1)))
            flow, flow_pl = self._flow(prec, f"predict_flow{lvl + 1}", src, segs, want_f32=True, want_planes=not UPFLOW_DIRECT)
            self._upflow(prec, f"upsampled_flow{lvl + 1}_to_{lvl}", flow, flow_pl, cat.window(2))
            join()
            src, segs = cat.buf, cat.seg
        flow2, _ = self._flow(prec, "predict_flow2", src, segs, want_f32=True, want_planes=False)
        return flow2


def _dims(x_or_hw, k):
    return x_or_hw // k


class FlowNetC(_Net):
    def __init__(self, args=None, batchNorm=True, div_flow=20):
        super().__init__()
        self.batchNorm, self.div_flow = batchNorm, div_flow
        self.conv1 = conv(batchNorm, 3, 64, kernel_size=7, stride=2)
        self.conv2 = conv(batchNorm, 64, 128, kernel_size=5, stride=2)
        self.conv3 = conv(batchNorm, 128, 256, kernel_size=5, stride=2)
        self.conv_redir = conv(batchNorm, 256, 32, kernel_size=1, stride=1)
        self.corr = Correlation(pad_size=20, kernel_size=1, max_displacement=20, stride1=1, stride2=2, corr_multiply=1)
        self.corr_activation = nn.LeakyReLU(0.1, inplace=True)
        self.conv3_1 = conv(batchNorm, 473, 256)
        self.conv4 = conv(batchNorm, 256, 512, stride=2)
        self.conv4_1 = conv(batchNorm, 512, 512)
        self.conv5 = conv(batchNorm, 512, 512, stride=2)
        self.conv5_1 = conv(batchNorm, 512, 512)
        self.conv6 = conv(batchNorm, 512, 1024, stride=2)
        self.conv6_1 = conv(batchNorm, 1024, 1024)
        self.deconv5 = deconv(1024, 512)
        self.deconv4 = deconv(1026, 256)
        self.deconv3 = deconv(770, 128)
        self.deconv2 = deconv(386, 64)
        self.predict_flow6 = predict_flow(1024)
        self.predict_flow5 = predict_flow(1026)
        self.predict_flow4 = predict_flow(770)
        self.predict_flow3 = predict_flow(386)
        self.predict_flow2 = predict_flow(194)
        self.upsampled_flow6_to_5 = nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=True)
        self.upsampled_flow5_to_4 = nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=True)
        self.upsampled_flow4_to_3 = nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=True)
        self.upsampled_flow3_to_2 = nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=True)
        _init(self)
        self.upsample1 = nn.Upsample(scale_factor=4, mode="bilinear")

    def run(self, x, prec):
        """x: f32 NCHW [B,6,H,W] -> flow2 f32 NHWC [B,H/4,W/4,2] (FlowNetC.py:71-128, eval)."""
        B, _, H, W = x.shape
        dev = x.device
        mk = lambda k, segs: self._concat(B, H // k, W // k, segs, prec, dev)
        cat5, cat4, cat3, cat2 = mk(32, (512, 512, 2)), mk(16, (512, 256, 2)), mk(8, (256, 128, 2)), mk(4, (128, 64, 2))
        x1, x2 = x[:, 0:3].contiguous(), x[:, 3:].contiguous()
        tower_b = {}
        join = self._fork(lambda: tower_b.update(c3=self._c(prec, "conv3", self._c(prec, "conv2", self._stem(prec, "conv1", x2)))),
                          allocates=True)  # the second frame's tower next to the first's
        c2a = self._c(prec, "conv2", self._stem(prec, "conv1", x1), out=cat2.window(0))
        c3a = self._c(prec, "conv3", c2a)
        join()
        c3b = tower_b["c3"]
        in31 = self._concat(B, H // 8, W // 8, (32, 441), prec, dev)
        # tensor-core cost volume straight from the planes; corr_activation and the layout of conv3_1's input in the gather
        ops.correlation_planes(c3a, c3b, 256, 20, 20, 2, out_planes=in31.window(1), act="leaky", act_param=LEAK)
        self._c(prec, "conv_redir", c3a, out=in31.window(0))
        c31 = self._c(prec, "conv3_1", in31.buf, out=cat3.window(0), segs=in31.seg)
        c4 = self._c(prec, "conv4_1", self._c(prec, "conv4", c31), out=cat4.window(0))
        c5 = self._c(prec, "conv5_1", self._c(prec, "conv5", c4), out=cat5.window(0))
        c6 = self._c(prec, "conv6_1", self._c(prec, "conv6", c5))
        return self._refine(prec, c6, (cat5, cat4, cat3, cat2), None)


class FlowNetS(_Net):
    def __init__(self, args=None, input_channels=12, batchNorm=True):
        super().__init__()
        self.batchNorm = batchNorm
        self.conv1 = conv(batchNorm, input_channels, 64, kernel_size=7, stride=2)
        self.conv2 = conv(batchNorm, 64, 128, kernel_size=5, stride=2)
        self.conv3 = conv(batchNorm, 128, 256, kernel_size=5, stride=2)
        self.conv3_1 = conv(batchNorm, 256, 256)
        self.conv4 = conv(batchNorm, 256, 512, stride=2)
        self.conv4_1 = conv(batchNorm, 512, 512)
        self.conv5 = conv(batchNorm, 512, 512, stride=2)
        self.conv5_1 = conv(batchNorm, 512, 512)
        self.conv6 = conv(batchNorm, 512, 1024, stride=2)
        self.conv6_1 = conv(batchNorm, 1024, 1024)
        self.deconv5 = deconv(1024, 512)
        self.deconv4 = deconv(1026, 256)
        self.deconv3 = deconv(770, 128)
        self.deconv2 = deconv(386, 64)
        self.predict_flow6 = predict_flow(1024)
        self.predict_flow5 = predict_flow(1026)
        self.predict_flow4 = predict_flow(770)
        self.predict_flow3 = predict_flow(386)
        self.predict_flow2 = predict_flow(194)
        self.upsampled_flow6_to_5 = nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=False)
        self.upsampled_flow5_to_4 = nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=False)
        self.upsampled_flow4_to_3 = nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=False)
        self.upsampled_flow3_to_2 = nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=False)
        _init(self)
        self.upsample1 = nn.Upsample(scale_factor=4, mode="bilinear")

    def run(self, x, prec):
        """x: f32 NCHW [B,12,H,W] -> flow2 f32 NHWC [B,H/4,W/4,2] (FlowNetS.py:60-94, eval)."""
        B, _, H, W = x.shape
        dev = x.device
        mk = lambda k, segs: self._concat(B, H // k, W // k, segs, prec, dev)
        cat5, cat4, cat3, cat2 = mk(32, (512, 512, 2)), mk(16, (512, 256, 2)), mk(8, (256, 128, 2)), mk(4, (128, 64, 2))
        c2 = self._c(prec, "conv2", self._stem(prec, "conv1", x), out=cat2.window(0))
        c3 = self._c(prec, "conv3_1", self._c(prec, "conv3", c2), out=cat3.window(0))
        c4 = self._c(prec, "conv4_1", self._c(prec, "conv4", c3), out=cat4.window(0))
        c5 = self._c(prec, "conv5_1", self._c(prec, "conv5", c4), out=cat5.window(0))
        c6 = self._c(prec, "conv6_1", self._c(prec, "conv6", c5))
        return self._refine(prec, c6, (cat5, cat4, cat3, cat2), None)


class FlowNetSD(_Net):
    def __init__(self, args=None, batchNorm=True):
        super().__init__()
        self.batchNorm = batchNorm
        self.conv0 = conv(batchNorm, 6, 64)
        self.conv1 = conv(batchNorm, 64, 64, stride=2)
        self.conv1_1 = conv(batchNorm, 64, 128)
        self.conv2 = conv(batchNorm, 128, 128, stride=2)
        self.conv2_1 = conv(batchNorm, 128, 128)
        self.conv3 = conv(batchNorm, 128, 256, stride=2)
        self.conv3_1 = conv(batchNorm, 256, 256)
        self.conv4 = conv(batchNorm, 256, 512, stride=2)
        self.conv4_1 = conv(batchNorm, 512, 512)
        self.conv5 = conv(batchNorm, 512, 512, stride=2)
        self.conv5_1 = conv(batchNorm, 512, 512)
        self.conv6 = conv(batchNorm, 512, 1024, stride=2)
        self.conv6_1 = conv(batchNorm, 1024, 1024)
        self.deconv5 = deconv(1024, 512)
        self.deconv4 = deconv(1026, 256)
        self.deconv3 = deconv(770, 128)
        self.deconv2 = deconv(386, 64)
        self.inter_conv5 = i_conv(batchNorm, 1026, 512)
        self.inter_conv4 = i_conv(batchNorm, 770, 256)
        self.inter_conv3 = i_conv(batchNorm, 386, 128)
        self.inter_conv2 = i_conv(batchNorm, 194, 64)
        self.predict_flow6 = predict_flow(1024)
        self.predict_flow5 = predict_flow(512)
        self.predict_flow4 = predict_flow(256)
        self.predict_flow3 = predict_flow(128)
        self.predict_flow2 = predict_flow(64)
        self.upsampled_flow6_to_5 = nn.ConvTranspose2d(2, 2, 4, 2, 1)
        self.upsampled_flow5_to_4 = nn.ConvTranspose2d(2, 2, 4, 2, 1)
        self.upsampled_flow4_to_3 = nn.ConvTranspose2d(2, 2, 4, 2, 1)
        self.upsampled_flow3_to_2 = nn.ConvTranspose2d(2, 2, 4, 2, 1)
        _init(self)
        self.upsample1 = nn.Upsample(scale_factor=4, mode="bilinear")

    def run(self, x, prec):
        """x: f32 NCHW [B,6,H,W] -> flow2 f32 NHWC [B,H/4,W/4,2] (FlowNetSD.py:66-106, eval)."""
        B, _, H, W = x.shape
        dev = x.device
        mk = lambda k, segs: self._concat(B, H // k, W // k, segs, prec, dev)
        cat5, cat4, cat3, cat2 = mk(32, (512, 512, 2)), mk(16, (512, 256, 2)), mk(8, (256, 128, 2)), mk(4, (128, 64, 2))
        c0 = self._stem(prec, "conv0", x)
        c1 = self._c(prec, "conv1_1", self._c(prec, "conv1", c0))
        c2 = self._c(prec, "conv2_1", self._c(prec, "conv2", c1), out=cat2.window(0))
        c3 = self._c(prec, "conv3_1", self._c(prec, "conv3", c2), out=cat3.window(0))
        c4 = self._c(prec, "conv4_1", self._c(prec, "conv4", c3), out=cat4.window(0))
        c5 = self._c(prec, "conv5_1", self._c(prec, "conv5", c4), out=cat5.window(0))
        c6 = self._c(prec, "conv6_1", self._c(prec, "conv6", c5))
        flow, flow_pl = self._flow(prec, "predict_flow6", c6, want_f32=True, want_planes=not UPFLOW_DIRECT)
        src, segs = c6, None
        for lvl, cat in zip((5, 4, 3, 2), (cat5, cat4, cat3, cat2)):
            join = self._fork(lambda: self._pc(prec, f"deconv{lvl}", segs)(src, post_act="leaky", act_param=LEAK,
                                                                          out_planes=cat.window(1)))
            self._upflow(prec, f"upsampled_flow{lvl + 1}_to_{lvl}", flow, flow_pl, cat.window(2))
            join()
            inter = self._c(prec, f"inter_conv{lvl}", cat.buf, segs=cat.seg, act=False)
            flow, flow_pl = self._flow(prec, f"predict_flow{lvl}", inter, want_f32=True,
                                       want_planes=lvl != 2 and not UPFLOW_DIRECT)
            src, segs = cat.buf, cat.seg
        return flow


class FlowNetFusion(_Net):
    def __init__(self, args=None, batchNorm=True):
        super().__init__()
        self.batchNorm = batchNorm
        self.conv0 = conv(batchNorm, 11, 64)
        self.conv1 = conv(batchNorm, 64, 64, stride=2)
        self.conv1_1 = conv(batchNorm, 64, 128)
        self.conv2 = conv(batchNorm, 128, 128, stride=2)
        self.conv2_1 = conv(batchNorm, 128, 128)
        self.deconv1 = deconv(128, 32)
        self.deconv0 = deconv(162, 16)
        self.inter_conv1 = i_conv(batchNorm, 162, 32)
        self.inter_conv0 = i_conv(batchNorm, 82, 16)
        self.predict_flow2 = predict_flow(128)
        self.predict_flow1 = predict_flow(32)
        self.predict_flow0 = predict_flow(16)
        self.upsampled_flow2_to_1 = nn.ConvTranspose2d(2, 2, 4, 2, 1)
        self.upsampled_flow1_to_0 = nn.ConvTranspose2d(2, 2, 4, 2, 1)
        _init(self)

    def run(self, x, prec):
        """x: f32 NCHW [B,11,H,W] -> flow0 f32 NHWC [B,H,W,2] (FlowNetFusion.py:47-67)."""
        B, _, H, W = x.shape
        dev = x.device
        cat1 = self._concat(B, H // 2, W // 2, (128, 32, 2), prec, dev, align=FUSION_CONCAT_ALIGN)
        cat0 = self._concat(B, H, W, (64, 16, 2), prec, dev, align=FUSION_CONCAT_ALIGN)
        c0 = self._stem(prec, "conv0", x, out=cat0.window(0))
        c1 = self._c(prec, "conv1_1", self._c(prec, "conv1", c0), out=cat1.window(0))
        c2 = self._c(prec, "conv2_1", self._c(prec, "conv2", c1))
        join = self._fork(lambda: self._pc(prec, "deconv1")(c2, post_act="leaky", act_param=LEAK, out_planes=cat1.window(1)))
        flow2, flow2_pl = self._flow(prec, "predict_flow2", c2, want_f32=True, want_planes=not UPFLOW_DIRECT)
        self._upflow(prec, "upsampled_flow2_to_1", flow2, flow2_pl, cat1.window(2))
        join()
        i1 = self._c(prec, "inter_conv1", cat1.buf, segs=cat1.seg, act=False)
        join = self._fork(lambda: self._pc(prec, "deconv0", cat1.seg)(cat1.buf, post_act="leaky", act_param=LEAK,
                                                                       out_planes=cat0.window(1)))
        flow1, flow1_pl = self._flow(prec, "predict_flow1", i1, want_f32=True, want_planes=not UPFLOW_DIRECT)
        self._upflow(prec, "upsampled_flow1_to_0", flow1, flow1_pl, cat0.window(2))
        join()
        i0 = self._c(prec, "inter_conv0", cat0.buf, segs=cat0.seg, act=False)
        flow0, _ = self._flow(prec, "predict_flow0", i0, want_f32=True, want_planes=False)
        return flow0


class MyDict(dict):
    pass


class FlowNet2(nn.Module):
    def __init__(self, args=None, batchNorm=False, div_flow=20.0):
        super().__init__()
        if args is None:
            args = MyDict()
            args.rgb_max = 1
            args.fp16 = False
            args.grads = {}
        self.batchNorm, self.div_flow, self.rgb_max, self.args = batchNorm, div_flow, args.rgb_max, args
        self.channelnorm = ChannelNorm()
        self.flownetc = FlowNetC(args, batchNorm=batchNorm)
        self.upsample1 = nn.Upsample(scale_factor=4, mode="bilinear")
        self.resample1 = Resample2d()
        self.flownets_1 = FlowNetS(args, batchNorm=batchNorm)
        self.upsample2 = nn.Upsample(scale_factor=4, mode="bilinear")
        self.resample2 = Resample2d()
        self.flownets_2 = FlowNetS(args, batchNorm=batchNorm)
        self.flownets_d = FlowNetSD(args, batchNorm=batchNorm)
        self.upsample3 = nn.Upsample(scale_factor=4, mode="nearest")
        self.upsample4 = nn.Upsample(scale_factor=4, mode="nearest")
        self.resample3 = Resample2d()
        self.resample4 = Resample2d()
        self.flownetfusion = FlowNetFusion(args, batchNorm=batchNorm)
        _init(self)
        self.precision = None
        self.parallel_sd = os.environ.get("SHINEON_FLOW_PARALLEL_SD", "1") != "0"
        self._side = {}

    def _side_stream(self, device):
        st = self._side.get(device)
        if st is None:
            st = self._side[device] = torch.cuda.Stream(device=device)
        return st

    def forward(self, inputs):
        """inputs f32 [B,3,2,H,W] (H, W multiples of 64) -> flow f32 NCHW [B,2,H,W] (models.py:127-192)."""
        require_cuda(self, "FlowNet2")
        if self.training:
            raise NotImplementedError("FlowNet2 training is not part of this build (the reference never trains it either)")
        prec = ops.resolve_precision(self.precision)
        df = self.div_flow
        x = ops.flownet_normalize(inputs.contiguous(), self.rgb_max)
        # FlowNetSD only needs the normalised pair: it runs on a side stream next to the C -> S1 -> S2 chain (most layers
        # of a try-on sized batch launch fewer CTAs than there are SMs) and joins before the fusion network.  Under
        # stream capture the fork / join becomes a parallel branch of the graph.
        cur = torch.cuda.current_stream(x.device)
        side = self._side_stream(x.device) if self.parallel_sd else None
        if side is not None:
            capturing = torch.cuda.is_current_stream_capturing()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                sd_flow = ops.upsample4x_flow(self.flownets_d.run(x, prec), 1.0 / df, bilinear=False)
            if not capturing:  # inside a capture the graph's private pool is not reused before the join below
                x.record_stream(side)
                sd_flow.record_stream(cur)
        c_flow = ops.upsample4x_flow(self.flownetc.run(x, prec), df, bilinear=True)
        s1_flow = ops.upsample4x_flow(self.flownets_1.run(ops.flownet_warp_concat(x, c_flow, df), prec), df, bilinear=True)
        s2_flow = ops.upsample4x_flow(self.flownets_2.run(ops.flownet_warp_concat(x, s1_flow, df), prec), df, bilinear=False)
        if side is not None:
            cur.wait_stream(side)
        else:
            sd_flow = ops.upsample4x_flow(self.flownets_d.run(x, prec), 1.0 / df, bilinear=False)
        flow0 = self.flownetfusion.run(ops.flownet_fusion_concat(x, sd_flow, s2_flow), prec)
        return flow0.permute(0, 3, 1, 2).contiguous()
