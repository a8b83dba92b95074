"""SelfAttention — SAGAN self-attention layer (reference: models/networks/attention/sagan.py:5-53).

Same ctor signature, sub-module names (query_conv / key_conv / value_conv / gamma) and state_dict keys.
Forward = one fused 1x1 tcgen05 conv for q|k|v + the sagan_attention kernel.
"""
import torch
from torch import nn

from ... import ops
from .._engine_util import params_signature, require_cuda


class SelfAttention(nn.Module):
    def __init__(self, in_dim, activation=nn.LeakyReLU):
        super().__init__()
        self.chanel_in = in_dim
        self.activation = activation
        self.query_conv = nn.Conv2d(in_channels=in_dim, out_channels=in_dim // 8, kernel_size=1)
        self.key_conv = nn.Conv2d(in_channels=in_dim, out_channels=in_dim // 8, kernel_size=1)
        self.value_conv = nn.Conv2d(in_channels=in_dim, out_channels=in_dim, kernel_size=1)
        self.gamma = nn.Parameter(torch.zeros(1))
        self.softmax = nn.Softmax(dim=-1)
        self._packed = None
        self.precision = None  # ops.PRECISIONS name; None = default (fp16x3)

    # ---- engine
    def packed(self, prec):
        sig = (params_signature(self), prec)
        if self._packed is None or self._packed[0] != sig:
            require_cuda(self, "SelfAttention")
            w = torch.cat([self.query_conv.weight, self.key_conv.weight, self.value_conv.weight], 0)
            b = torch.cat([self.query_conv.bias, self.key_conv.bias, self.value_conv.bias], 0)
            self._packed = (sig, ops.PackedConv(w, b, stride=1, pad=0, prec=prec))
        return self._packed[1]

    def run(self, x_f32, x_planes, *, act=None, act_param=0.0, want_f32=False, want_planes=True):
        """x_f32: f32 NHWC [N,H,W,C]; x_planes: the same values as planes.  Returns (f32|None, Planes|None) of
        act(gamma * attention(x) + x)."""
        prec = x_planes.prec
        qkv, _ = ops.conv2d(x_planes, self.packed(prec), want_f32=True)
        return ops.sagan_attention(qkv, x_f32, self.gamma.detach(), self.chanel_in // 8, act=act, act_param=act_param,
                                   want_f32=want_f32, want_planes=want_planes, prec=prec)

    def forward(self, x):
        """x: [B, C, W, H] f32 CUDA tensor -> gamma * attention(x) + x (sagan.py:29-53)."""
        xp = ops.nchw_to_planes(x.contiguous(), prec=ops.resolve_precision(self.precision))
        x_nhwc = x.permute(0, 2, 3, 1).contiguous()
        y, _ = self.run(x_nhwc, xp, want_f32=True, want_planes=False)
        return y.permute(0, 3, 1, 2).contiguous()
