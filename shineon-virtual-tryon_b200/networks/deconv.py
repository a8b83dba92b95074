"""ConvTranspose2d(kernel 4, stride 2, padding 1) on the tcgen05 conv kernel.

Reference: `deconv` / `upsampled_flow*` in models/flownet2_pytorch/networks/submodules.py:34-38 and
FlowNetC.py:59-62.  out[2m+py] = sum over the two taps of that output parity, i.e. four phase-wise
2x2 stride-1 convolutions of the *input*, each scattered to its (py, px) output lattice:
  py = 0: x[m-1]*w[3] + x[m]*w[1]   (pad 1 before)      py = 1: x[m]*w[2] + x[m+1]*w[0]   (pad 0)
"""
import torch

from .. import ops

_TAPS = {0: [3, 1], 1: [2, 0]}


class PackedDeconv4x4s2:
    def __init__(self, weight, bias=None, cin_pad=None, prec=None, chan_map=None):
        # weight: [Cin, Cout, 4, 4] (nn.ConvTranspose2d layout)
        assert weight.shape[2] == 4 and weight.shape[3] == 4
        self.Cout = weight.shape[1]
        w = weight.detach().float().permute(1, 0, 2, 3)  # -> [Cout, Cin, 4, 4]
        self.phases = []
        for py in (0, 1):
            for px in (0, 1):
                wk = w[:, :, _TAPS[py]][:, :, :, _TAPS[px]].contiguous()
                pc = ops.PackedConv(wk, bias, stride=1, pad_hw=(1 - py, 1 - px), cin_pad=cin_pad, prec=prec,
                                    chan_map=chan_map)
                self.phases.append((py, px, pc))

    def __call__(self, x, *, post_act=None, act_param=0.0, want_f32=False, want_planes=False, out_f32=None,
                 out_planes=None, out_coffset=0):
        N, H, W = x.N, x.H, x.W
        dev = x.hi.device
        if want_f32 and out_f32 is None:
            out_f32 = torch.empty(N, 2 * H, 2 * W, self.Cout, dtype=torch.float32, device=dev)
        if want_planes and out_planes is None:
            out_planes = ops.Planes(N, 2 * H, 2 * W, self.Cout, prec=x.prec, device=dev)
        for py, px, pc in self.phases:
            ops.conv2d(x, pc, post_act=post_act, act_param=act_param, out_f32=out_f32, out_planes=out_planes,
                       out_coffset=out_coffset, out_geom=(2 * H, 2 * W, 2, py, 2, px), out_hw=(H, W))
        return out_f32, out_planes
