"""AttentiveMultiSpade: the SPADEs run side by side on the same input, their outputs are stacked on channels, passed
through SAGAN self-attention and reduced by a conv + LeakyReLU (reference: models/networks/sams/attentive_multispade.py:11-50)."""
import torch
from torch import nn

from ... import ops
from .._engine_util import params_signature, require_cuda
from ..attention import ATTENTION_TYPES
from .multispade import MultiSpade
from .spade import SPADE


class AttentiveMultiSpade(MultiSpade):
    def __init__(self, config_text, norm_nc, label_channels_dict, activation, attn_type="sagan"):
        super().__init__(config_text, norm_nc, label_channels_dict, activation)
        _, kernel_size = SPADE.parse_config_text(config_text)
        self.attn_nc = norm_nc * len(self.spade_layers)
        self.attention_layer = ATTENTION_TYPES[attn_type](self.attn_nc)
        self.mlp_final = nn.Sequential(nn.Conv2d(self.attn_nc, norm_nc, kernel_size=kernel_size, padding=kernel_size // 2),
                                       nn.LeakyReLU())
        self._packed_final = None

    def _final(self, prec):
        sig = (params_signature(self.mlp_final), prec)
        if self._packed_final is None or self._packed_final[0] != sig:
            require_cuda(self, "AttentiveMultiSpade")
            c = self.mlp_final[0]
            self._packed_final = (sig, ops.PackedConv(c.weight, c.bias, stride=1, pad=c.padding[0], prec=prec))
        return self._packed_final[1]

    def run(self, ctx, x, labelmap_dict, *, act=None, act_param=0.0, shared=None, want_f32=False, want_planes=True, out_f32=None,
            out_planes=None):
        items = self._ordered(labelmap_dict)
        N, H, W, C = x.shape
        stack = torch.empty(N, H, W, self.attn_nc, dtype=torch.float32, device=x.device)
        stack_p = ops.Planes(N, H, W, self.attn_nc, prec=ctx.prec, device=x.device)
        for i, (key, seg) in enumerate(items):  # attentive_multispade.py:37-43: channel order = sorted key order
            win = stack_p.window(0, self.attn_nc)
            win.coffset, win.C = stack_p.coffset + i * C, C  # modulate's stores only need 8-byte alignment
            self.spade_layers[key].run(ctx, x, seg, shared=shared, out_f32=stack, out_f32_coffset=i * C, out_planes=win)
        _, attended = self.attention_layer.run(stack, stack_p, want_f32=False, want_planes=True)
        slope = float(self.mlp_final[1].negative_slope)
        if act is None:
            return ops.conv2d(attended, self._final(ctx.prec), post_act="leaky", act_param=slope, want_f32=want_f32,
                              want_planes=want_planes, out_f32=out_f32, out_planes=out_planes)
        # the block's activation follows mlp_final's LeakyReLU: one more pointwise pass over the (small) attended levels
        y, _ = ops.conv2d(attended, self._final(ctx.prec), post_act="leaky", act_param=slope, want_f32=True)
        return ops.instnorm_act(y, do_norm=False, act=act, act_param=act_param, want_f32=want_f32, want_planes=want_planes,
                                prec=ctx.prec, out_f32=out_f32, out_planes=out_planes)
