"""SelfAttention — SAGAN self-attention layer (reference: models/networks/attention/sagan.py:5-53).

Same ctor signature, sub-module names (query_conv / key_conv / value_conv / gamma) and state_dict keys.
Forward = one fused 1x1 tcgen05 conv for q|k|v + the sagan_attention kernel.
"""
import torch
from torch import nn

from ... import ops
from .._engine_util import params_signature, require_cuda, side_run


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

    def run(self, x_f32, x_planes, *, act=None, act_param=0.0, want_f32=False, want_planes=True, out_planes=None):
        """x_f32: f32 NHWC [N,H,W,C]; x_planes: the same values as planes.  Returns (f32|None, Planes|None) of
        act(gamma * attention(x) + x)."""
        prec = x_planes.prec
        qkv, _ = ops.conv2d(x_planes, self.packed(prec), want_f32=True)
        return ops.sagan_attention(qkv, x_f32, self.gamma.detach(), self.chanel_in // 8, act=act, act_param=act_param,
                                   want_f32=want_f32, want_planes=want_planes, prec=prec, out_planes=out_planes)

    # ---- training (row U6)
    def run_train(self, x_f32, x_planes):
        """-> (z = gamma*attention(x) + x as f32 NHWC, qkv projections f32 NHWC) for the backward."""
        qkv, _ = ops.conv2d(x_planes, self.packed(x_planes.prec), want_f32=True)
        z, _ = ops.sagan_attention(qkv, x_f32, self.gamma.detach(), self.chanel_in // 8, want_f32=True, want_planes=False)
        return z, qkv

    def packed_dgrad(self, prec):
        """The stacked q/k/v projection weights as the input-gradient conv's operand (re-packed when the weights change)."""
        sig = (params_signature(self), prec, "dgrad")
        if getattr(self, "_packed_dgrad", None) is None or self._packed_dgrad[0] != sig:
            w = torch.cat([self.query_conv.weight, self.key_conv.weight, self.value_conv.weight], 0)
            self._packed_dgrad = (sig, ops.PackedConv(w, None, stride=1, pad=0, prec=prec, transposed=True))
        return self._packed_dgrad[1]

    def backward(self, x_planes, qkv, g_out, prec):
        """g_out = dL/dz (f32 NHWC).  Accumulates gamma / q,k,v weight and bias gradients; returns the part of dL/dx that
        flows through the projections (the residual branch contributes g_out itself, added by the caller)."""
        from ..cpvton.unet import _grad_of

        C, Cq = self.chanel_in, self.chanel_in // 8
        g_qkv = ops.sagan_attention_bwd(qkv, self.gamma.detach(), g_out, Cq, _grad_of(self.gamma), beta_gamma=1.0)
        _, G = ops.instnorm_act(g_qkv, do_norm=False, want_f32=False, want_planes=True, prec=prec)  # f32 -> planes
        def proj_wgrad():  # weight / bias gradients of the three projections: next to the input-gradient conv below
            for conv, c0, cn in ((self.query_conv, 0, Cq), (self.key_conv, Cq, Cq), (self.value_conv, 2 * Cq, C)):
                ops.channel_sum(g_qkv, _grad_of(conv.bias), beta=1.0, coffset=c0)
                ops.conv2d_wgrad(G, x_planes, _grad_of(conv.weight), Cout=cn, Cin=C, kh=1, kw=1, stride=1, pad=0, beta=1.0,
                                 g_coffset=c0)

        from ..cpvton import unet as _unet

        join = side_run(proj_wgrad, _unet.GRAD_READY_HOOK is None)
        g_x, _ = ops.conv2d(G, self.packed_dgrad(prec), want_f32=True)
        join()  # g_qkv / G / x_planes stay referenced until here
        return g_x

    def forward(self, x):
        """x: [B, C, W, H] f32 CUDA tensor -> gamma * attention(x) + x (sagan.py:29-53)."""
        xp = ops.nchw_to_planes(x.contiguous(), prec=ops.resolve_precision(self.precision))
        x_nhwc = x.permute(0, 2, 3, 1).contiguous()
        y, _ = self.run(x_nhwc, xp, want_f32=True, want_planes=False)
        return y.permute(0, 3, 1, 2).contiguous()
