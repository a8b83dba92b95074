"""SamsGenerator: the Self-Attentive Multi-SPADE generator (reference: models/networks/sams/sams_generator.py:19-317).

Encoder: Conv3x3 over the previous frames, then [SPADE ResBlock -> nearest x0.5] per power step; middle: MultiSpade /
AttentiveMultiSpade ResBlocks at the innermost width; decoder: [nearest x2 -> ResBlock] per power step, Conv3x3 to RGB
(+ 1 mask channel with flow_warp).  The module tree equals the reference's (nn.ModuleList indices included, so checkpoints
load with strict=True); the forward walks it once and drives the kernels with activations as f32 NHWC between blocks.
"""
import argparse
import logging
import sys

import torch
from torch import nn

from ... import ops
from ...models.base_model import CHANNELS
from .._engine_util import params_signature, require_cuda
from .attentive_multispade import AttentiveMultiSpade
from .multispade import MultiSpade
from .spade import SPADE, AnySpadeResBlock, EngineContext

logger = logging.getLogger("logger")


def _label_nc(name):
    return CHANNELS[name.upper()]


def _int_list(v):
    """--attention_*_indices arrive as a list of strings, one string, or nothing (nargs='?')."""
    if v is None:
        return []
    if isinstance(v, str):
        return [t for t in v.replace(",", " ").split() if t]
    return [str(t) for t in v]


def choose_spade_class_by_index(attn_indices, i, total_layers):
    return AttentiveMultiSpade if (str(i) in attn_indices or str(i - total_layers) in attn_indices) else MultiSpade


def make_encode_block(in_feat, out_feat, **spade_kwargs):
    return [AnySpadeResBlock(in_feat, out_feat, **spade_kwargs), nn.Upsample(scale_factor=0.5)]


def make_decode_block(in_feat, out_feat, **spade_kwargs):
    return [nn.Upsample(scale_factor=2), AnySpadeResBlock(in_feat, out_feat, **spade_kwargs)]


class SamsGenerator(nn.Module):
    @classmethod
    def modify_commandline_options(cls, parser: argparse.ArgumentParser, is_train):
        parser.add_argument("--norm_G", default="spectralspadesyncbatch3x3")
        parser.add_argument("--ngf_base", type=int, default=2, help="Control the size of the network. ngf_base ** pow")
        parser.add_argument("--ngf_power_start", "--ngf_pow_outer", dest="ngf_pow_outer", type=int, default=6,
                            help="number of features at the outer ends = ngf_base ** ngf_pow_outer")
        parser.add_argument("--ngf_power_end", "--ngf_pow_inner", dest="ngf_pow_inner", type=int, default=10,
                            help="INCLUSIVE: features in the middle of the network = ngf_base ** ngf_pow_inner")
        parser.add_argument("--ngf_pow_step", type=int, default=1, help="power increment between layers")
        parser.add_argument("--num_middle", type=int, default=3, help="channel-preserving layers between encoder and decoder")
        parser.add_argument("--attention_middle_indices", nargs="?", default=[], help="middle layer indices for attention")
        parser.add_argument("--attention_decoder_indices", nargs="?", default=[], help="decoder layer indices for attention")
        if "--ngf" in sys.argv:
            logger.warning("SamsGenerator does NOT use --ngf. Use --ngf_base, --ngf_pow_outer, --ngf_pow_inner, "
                           "--ngf_pow_step, and --num_middle to control the architecture.")
        return parser

    def __init__(self, hparams):
        super().__init__()
        assert hparams.ngf_base > 1, f"{hparams.ngf_base}"
        assert hparams.ngf_pow_inner >= 1, f"{hparams.ngf_pow_inner=}"
        self.hparams = hparams
        self.inputs = hparams.person_inputs + hparams.cloth_inputs
        num_prev = max(hparams.n_frames_total - 1, 1)
        self.in_channels = 3 * num_prev
        out_channels = 3 + (1 if hparams.flow_warp else 0)  # + the flow-blend weight mask
        base, p_out, p_in, step = hparams.ngf_base, hparams.ngf_pow_outer, hparams.ngf_pow_inner, hparams.ngf_pow_step
        outer, inner = int(base ** p_out), int(base ** p_in)
        out_feat = outer

        kwargs = dict(norm_G=hparams.norm_G, label_channels=_label_nc(hparams.encoder_input) * num_prev,
                      activation=hparams.activation)
        enc = [nn.Conv2d(self.in_channels, outer, kernel_size=3, padding=1)]
        for p in range(p_out, p_in, step):
            out_feat = int(base ** (p + step))
            enc.extend(make_encode_block(int(base ** p), out_feat, **kwargs, spade_class=SPADE))
        if out_feat != inner:  # the power range did not land on the innermost width: one more block
            enc.extend(make_encode_block(out_feat, inner, **kwargs, spade_class=SPADE))
        self.encode_layers = nn.ModuleList(enc)

        kwargs["label_channels"] = {inp: _label_nc(inp) for inp in sorted(self.inputs)}
        mid_attn = _int_list(hparams.attention_middle_indices)
        self.middle_layers = nn.ModuleList(
            AnySpadeResBlock(inner, inner, **kwargs, spade_class=choose_spade_class_by_index(mid_attn, i, hparams.num_middle))
            for i in range(hparams.num_middle))

        dec_attn = _int_list(hparams.attention_decoder_indices)
        pows = list(range(p_in, p_out, -step))
        dec = []
        for i, p in enumerate(pows):
            out_feat = int(base ** (p - step))
            dec.extend(make_decode_block(int(base ** p), out_feat, **kwargs,
                                         spade_class=choose_spade_class_by_index(dec_attn, i, len(pows))))
        if out_feat != outer:
            dec.extend(make_decode_block(out_feat, outer, **kwargs,
                                         spade_class=AttentiveMultiSpade if dec_attn else MultiSpade))
        dec.append(nn.Conv2d(outer, out_channels, kernel_size=3, padding=1))
        self.decode_layers = nn.ModuleList(dec)
        self.precision = None  # ops.PRECISIONS name; None = default fp16x3 (fp32-grade products)
        self._packed = None

    def _pack(self, prec):
        first, last = self.encode_layers[0], self.decode_layers[-1]
        sig = (params_signature(first), params_signature(last), prec)
        if self._packed is None or self._packed[0] != sig:
            pf = ops.PackedConv(first.weight, first.bias, stride=1, pad=1, prec=prec)
            if last.out_channels <= 8:  # few output channels: tap-stacked 1x1 GEMM + col2im
                pl = ops.TapStackedConv3x3(last.weight, last.bias, prec=prec)
            else:
                pl = ops.PackedConv(last.weight, last.bias, stride=1, pad=1, prec=prec)
            self._packed = (sig, (pf, pl))
        return self._packed[1]

    def forward_nhwc(self, prev_n_frames_G, prev_n_labelmaps, current_labelmap_dict):
        require_cuda(self, "SamsGenerator")
        if self.training:
            raise NotImplementedError("the native SAMS generator is an inference engine (spectral-norm power iterations and "
                                      "batch statistics are training-time host logic of the reference): call .eval()")
        hp = self.hparams
        prec = ops.resolve_precision(self.precision)
        maps = {k: v.contiguous() for k, v in current_labelmap_dict.items()}
        if hp.n_frames_total > 1:
            b, n, c, h, w = prev_n_frames_G.shape
            x_in = prev_n_frames_G.reshape(b, n * c, h, w).contiguous()
            prev_maps = prev_n_labelmaps.reshape(b, -1, h, w).contiguous()
        else:  # sams_generator.py:263-270: a zero previous frame / label map
            ref = next(iter(maps.values()))
            b, _, h, w = ref.shape
            x_in = torch.zeros(b, self.in_channels, h, w, dtype=torch.float32, device=ref.device)
            prev_maps = torch.zeros(b, _label_nc(hp.encoder_input), h, w, dtype=torch.float32, device=ref.device)
        pf, pl = self._pack(prec)
        ctx = EngineContext(prec)
        x, _ = ops.conv2d(ops.nchw_to_planes(x_in.float(), prec=prec), pf, want_f32=True)
        for layer in list(self.encode_layers)[1:]:
            x = layer.run(ctx, x, prev_maps) if isinstance(layer, AnySpadeResBlock) else ops.nearest_resize_nhwc(x, layer.scale_factor)
        for layer in self.middle_layers:
            x = layer.run(ctx, x, maps)
        for layer in list(self.decode_layers)[:-1]:
            x = layer.run(ctx, x, maps) if isinstance(layer, AnySpadeResBlock) else ops.nearest_resize_nhwc(x, layer.scale_factor)
        _, xp = ops.instnorm_act(x, do_norm=False, want_f32=False, want_planes=True, prec=prec)
        if isinstance(pl, ops.TapStackedConv3x3):
            return pl(xp)
        return ops.conv2d(xp, pl, want_f32=True)[0]

    def forward(self, prev_n_frames_G, prev_n_labelmaps, current_labelmap_dict):
        """sams_generator.py:241-292.  prev_*: [b, n-1, c, h, w] (None when n_frames_total == 1); current_labelmap_dict:
        {input name: [b, c, h, w]}.  Returns the synthesized frame [b, 3 (+1), h, w]."""
        return self.forward_nhwc(prev_n_frames_G, prev_n_labelmaps, current_labelmap_dict).permute(0, 3, 1, 2).contiguous()
