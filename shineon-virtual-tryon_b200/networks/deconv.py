"""ConvTranspose2d(kernel 4, stride 2, padding 1) on the tcgen05 conv kernel.

Reference: `deconv` / `upsampled_flow*` in models/flownet2_pytorch/networks/submodules.py:34-38 and
FlowNetC.py:59-62.  out[2m+py] = sum over the two taps of that output parity, i.e. four phase-wise
2x2 stride-1 convolutions of the *input*, each scattered to its (py, px) output lattice:
  py = 0: x[m-1]*w[3] + x[m]*w[1]   (pad 1 before)      py = 1: x[m]*w[2] + x[m+1]*w[0]   (pad 0)
"""
import os

import torch

from .. import ops

_TAPS = {0: [3, 1], 1: [2, 0]}
# True: one launch per transposed conv (phase = a tile dimension of conv_igemm); False: four phase launches (round 1 / 2a)
MERGE_PHASES = os.environ.get("SHINEON_DECONV_MERGE", "1") != "0"


class PackedDeconv4x4s2:
    def __init__(self, weight, bias=None, cin_pad=None, prec=None, chan_map=None):
        # weight: [Cin, Cout, 4, 4] (nn.ConvTranspose2d layout)
        assert weight.shape[2] == 4 and weight.shape[3] == 4
        Cin, Cout = weight.shape[0], weight.shape[1]
        self.Cout = Cout
        fmt, split = ops.resolve_precision(prec)
        weight = ops._req(weight.detach().float().contiguous(), name="weight")
        cin_pad = ops.cpad64(Cin) if cin_pad is None else cin_pad
        dev = weight.device
        w_hi = torch.empty(4, Cout, 4, cin_pad, dtype=ops._DTYPES[fmt], device=dev)
        w_lo = torch.empty_like(w_hi) if split else None
        w_scale = 1.0
        if fmt == ops._lib.FMT_FP16:  # keep hi/lo out of the fp16 subnormals (see ops.PackedConv)
            amax = float(weight.abs().max())
            if amax > 0:
                import math

                w_scale = 2.0 ** math.floor(math.log2(16384.0 / amax))
        cm = None
        if chan_map is not None:
            cm = torch.as_tensor(chan_map, dtype=torch.int32, device=dev).contiguous()
            assert cm.numel() == cin_pad
        ops.check(ops._lib.load().shineon_pack_deconv4x4s2_weight(ops._p(weight), ops._p(w_hi), ops._p(w_lo), Cin, Cout,
                                                                  cin_pad, ops._p(cm), fmt, w_scale, ops._stream()),
                  "shineon_pack_deconv4x4s2_weight")
        b = None if bias is None else ops._req(bias.detach().float().contiguous(), name="bias")
        # all four phases in one launch (the phase is a tile dimension of the kernel): pc_all holds the whole weight batch
        pa = ops.PackedConv.__new__(ops.PackedConv)
        pa.fmt, pa.Cout, pa.Cin, pa.kh, pa.kw, pa.stride = fmt, Cout, Cin, 2, 2, 1
        pa.pad_h, pa.pad_w, pa.cin_pad = 1, 1, cin_pad
        pa.w_hi, pa.w_lo = w_hi, w_lo
        pa.acc_scale, pa.bias, pa.transposed = 1.0 / w_scale, b, False
        self.pc_all = pa
        self.phases = []
        for py in (0, 1):
            for px in (0, 1):
                pc = ops.PackedConv.__new__(ops.PackedConv)
                pc.fmt, pc.Cout, pc.Cin, pc.kh, pc.kw, pc.stride = fmt, Cout, Cin, 2, 2, 1
                pc.pad_h, pc.pad_w, pc.cin_pad = 1 - py, 1 - px, cin_pad
                pc.w_hi, pc.w_lo = w_hi[py * 2 + px], (w_lo[py * 2 + px] if split else None)
                pc.acc_scale, pc.bias, pc.transposed = 1.0 / w_scale, b, False
                self.phases.append((py, px, pc))

    def __call__(self, x, *, post_act=None, act_param=0.0, want_f32=False, want_planes=False, out_f32=None,
                 out_planes=None, out_coffset=0):
        N, H, W = x.N, x.H, x.W
        dev = x.hi.device
        if want_f32 and out_f32 is None:
            out_f32 = torch.empty(N, 2 * H, 2 * W, self.Cout, dtype=torch.float32, device=dev)
        if want_planes and out_planes is None:
            out_planes = ops.Planes(N, 2 * H, 2 * W, self.Cout, prec=x.prec, device=dev)
        if MERGE_PHASES:
            ops.conv2d(x, self.pc_all, post_act=post_act, act_param=act_param, out_f32=out_f32, out_planes=out_planes,
                       out_coffset=out_coffset, out_geom=(2 * H, 2 * W, 2, 0, 2, 0), out_hw=(H, W), deconv_phases=True)
            return out_f32, out_planes
        for py, px, pc in self.phases:
            ops.conv2d(x, pc, post_act=post_act, act_param=act_param, out_f32=out_f32, out_planes=out_planes,
                       out_coffset=out_coffset, out_geom=(2 * H, 2 * W, 2, py, 2, px), out_hw=(H, W))
        return out_f32, out_planes
