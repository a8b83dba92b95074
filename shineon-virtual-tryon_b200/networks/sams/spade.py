"""SPADE and AnySpadeResBlock (reference: models/networks/sams/spade.py:17-183).

Engine view of one SPADE layer (spade.py:68-84), x kept as f32 NHWC between passes:
    label map --nearest resize + im2col (one pass, cached per map and resolution)--> operand planes, K = pad64(ks*ks*label_nc)
    operand --ONE dense GEMM for the mlp_shared convs of ALL SPADE layers of the block that read this map, + act--> actv planes
              (128 channels per layer, each layer reads its window)
    actv  --ONE conv ks x ks with mlp_gamma | mlp_beta stacked (bias of gamma + 1)-->  (1+gamma | beta) f32
    spade_modulate: param-free norm of x, * (1+gamma) + beta, the block's activation, hi/lo split  -->  conv operand
"""
import re

import torch
from torch import nn
from torch.nn.utils import spectral_norm

from ... import ops
from .._engine_util import params_signature, require_cuda
from ..activation import Sine, Swish, act_name


class SynchronizedBatchNorm2d(nn.BatchNorm2d):
    """models/networks/sync_batchnorm/batchnorm.py: outside DataParallel replication (and always in eval mode) the
    reference class IS F.batch_norm over its own running statistics (batchnorm.py:68-73); same buffers / keys."""


def _activation(activation, resblock):
    """spade.py:86-96 (SPADE: relu -> ReLU) and spade.py:173-183 (block: relu -> LeakyReLU(0.2))."""
    if activation == "relu":
        return nn.LeakyReLU(2e-1) if resblock else nn.ReLU()
    if activation == "gelu":
        return nn.GELU()
    if activation == "swish":
        return Swish()
    if activation == "sine":
        return Sine()
    raise RuntimeError(f"The selected activation should be relu/gelu/swish/sine, not {activation}")


def effective_weight(conv):
    """Eval-mode weight of a conv that may be wrapped by torch.nn.utils.spectral_norm: weight_orig / (u . W v), the
    value the reference's forward pre-hook computes without a power iteration when the module is not training."""
    if hasattr(conv, "weight_orig"):
        w0 = conv.weight_orig.detach()
        sigma = torch.dot(conv.weight_u, torch.mv(w0.reshape(w0.shape[0], -1), conv.weight_v))
        return w0 / sigma
    return conv.weight.detach()


def shared_gemm_weight(spades):
    """mlp_shared weights of `spades` (all reading the same label map) as ONE 1x1 GEMM weight [128 * len, ks*ks*label_nc, 1, 1]
    over the im2col operand (k = (fy*ks + fx)*label_nc + c, ops.nearest_im2col_planes)."""
    ws = [sp.mlp_shared[0].weight.detach().float() for sp in spades]
    return torch.cat([w.permute(0, 2, 3, 1).reshape(w.shape[0], -1, 1, 1) for w in ws], 0).contiguous()


class EngineContext:
    """State of one generator pass: numeric mode, label-map planes per resolution, InstanceNorm statistics per tensor."""

    def __init__(self, prec):
        self.prec = prec
        self._seg = {}
        self._stats = {}

    def seg_operand(self, seg, H, W, ks):
        """im2col planes (ks x ks, pad ks/2) of the label map nearest-resized to (H, W): the A operand of mlp_shared."""
        key = (seg.data_ptr(), H, W, ks)
        if key not in self._seg:
            self._seg[key] = (seg, ops.nearest_im2col_planes(seg, (H, W), ks, prec=self.prec))  # keeps `seg` alive: the key is its address
        return self._seg[key][1]

    def stats(self, x):
        key = x.data_ptr()
        if key not in self._stats:
            self._stats[key] = (x, ops.chan_stats(x))
        return self._stats[key][1]


class SPADE(nn.Module):
    @staticmethod
    def parse_config_text(config_text):
        assert config_text.startswith("spade")
        parsed = re.search(r"spade(\D+)(\d)x\d", config_text)
        kind = str(parsed.group(1))
        if kind == "instance":
            norm = nn.InstanceNorm2d
        elif kind == "syncbatch":
            norm = SynchronizedBatchNorm2d
        elif kind == "batch":
            norm = nn.BatchNorm2d
        else:
            raise ValueError("%s is not a recognized param-free norm type in SPADE" % kind)
        return norm, int(parsed.group(2))

    def __init__(self, config_text, norm_nc, label_nc, activation):
        super().__init__()
        norm_cls, ks = SPADE.parse_config_text(config_text)
        self.param_free_norm = norm_cls(norm_nc, affine=False)
        self.actvn = _activation(activation, resblock=False)
        self.nhidden = nhidden = 128
        pw = ks // 2
        self.mlp_shared = nn.Sequential(nn.Conv2d(label_nc, nhidden, kernel_size=ks, padding=pw), self.actvn)
        self.mlp_gamma = nn.Conv2d(nhidden, norm_nc, kernel_size=ks, padding=pw)
        self.mlp_beta = nn.Conv2d(nhidden, norm_nc, kernel_size=ks, padding=pw)
        self._packed = None

    # ---- engine
    def _pack(self, prec):
        sig = (params_signature(self), prec)
        if self._packed is not None and self._packed[0] == sig:
            return self._packed[1]
        require_cuda(self, "SPADE")
        sh = self.mlp_shared[0]
        pw = sh.padding[0]
        d = dict(shared=ops.PackedConv(shared_gemm_weight([self]), sh.bias, stride=1, pad=0, prec=prec))
        w = torch.cat([self.mlp_gamma.weight, self.mlp_beta.weight], 0)
        b = torch.cat([self.mlp_gamma.bias + 1.0, self.mlp_beta.bias], 0)  # (1 + gamma) | beta
        d["gb"] = ops.PackedConv(w, b, stride=1, pad=pw, prec=prec)
        norm = self.param_free_norm
        d["bn"] = None
        if isinstance(norm, nn.BatchNorm2d):
            rstd = torch.rsqrt(norm.running_var.detach().float() + norm.eps)
            d["bn"] = (rstd.contiguous(), (-norm.running_mean.detach().float() * rstd).contiguous())
        elif not isinstance(norm, nn.InstanceNorm2d):
            raise NotImplementedError(f"param-free norm {type(norm).__name__} has no native kernel")
        self._packed = (sig, d)
        return d

    def run(self, ctx, x, seg, *, act=None, act_param=0.0, actv=None, shared=None, **out):
        """x: f32 NHWC [N,H,W,norm_nc]; seg: f32 NCHW label map at any resolution.  actv: this layer's mlp_shared output if the
        caller already has it (AnySpadeResBlock computes the mlp_shared convs of all its layers per label map in one GEMM).
        `out`: the output selection of ops.spade_modulate (want_f32 / want_planes / out_f32 / out_f32_coffset / out_planes).
        Returns (f32|None, Planes|None)."""
        norm = self.param_free_norm
        if self.training and isinstance(norm, nn.BatchNorm2d):
            raise NotImplementedError("SPADE with batch statistics (training mode) has no native kernel; call .eval()")
        pk = self._pack(ctx.prec)
        N, H, W, _ = x.shape
        if actv is None and shared:
            actv = shared.get(id(self))
        if actv is None:
            a, ap = act_name(self.actvn)
            ks = self.mlp_shared[0].kernel_size[0]
            _, actv = ops.conv2d(ctx.seg_operand(seg, H, W, ks), pk["shared"], post_act=a, act_param=ap, want_planes=True)
        gb, _ = ops.conv2d(actv, pk["gb"], want_f32=True)
        if pk["bn"] is not None:
            nk = dict(nscale=pk["bn"][0], nshift=pk["bn"][1])
        else:
            nk = dict(stats_ws=ctx.stats(x))
        return ops.spade_modulate(x, gb, eps=norm.eps, act=act, act_param=act_param, prec=ctx.prec, **nk, **out)

    precision = None

    def forward(self, x, segmap):
        """spade.py:68-84 on the reference layout (f32 NCHW CUDA tensors)."""
        ctx = EngineContext(ops.resolve_precision(self.precision))
        y, _ = self.run(ctx, x.permute(0, 2, 3, 1).contiguous(), segmap.contiguous(), want_f32=True, want_planes=False)
        return y.permute(0, 3, 1, 2).contiguous()


class AnySpadeResBlock(nn.Module):
    def __init__(self, fin, fout, norm_G, label_channels, spade_class, activation):
        super().__init__()
        self.learned_shortcut = fin != fout
        fmiddle = min(fin, fout)
        self.conv_0 = nn.Conv2d(fin, fmiddle, kernel_size=3, padding=1)
        self.conv_1 = nn.Conv2d(fmiddle, fout, kernel_size=3, padding=1)
        if self.learned_shortcut:
            self.conv_s = nn.Conv2d(fin, fout, kernel_size=1, bias=False)
        if "spectral" in norm_G:
            self.conv_0 = spectral_norm(self.conv_0)
            self.conv_1 = spectral_norm(self.conv_1)
            if self.learned_shortcut:
                self.conv_s = spectral_norm(self.conv_s)
        cfg = norm_G.replace("spectral", "")
        self.spade_0 = spade_class(cfg, fin, label_channels, activation)
        self.spade_1 = spade_class(cfg, fmiddle, label_channels, activation)
        if self.learned_shortcut:
            self.norm_s = spade_class(cfg, fin, label_channels, activation)
        self.actvn = _activation(activation, resblock=True)
        self._packed = None

    def _convs(self):
        return [self.conv_0, self.conv_1] + ([self.conv_s] if self.learned_shortcut else [])

    def _pack(self, prec):
        sig = (tuple(params_signature(c) for c in self._convs()), prec)
        if self._packed is not None and self._packed[0] == sig:
            return self._packed[1]
        require_cuda(self, "AnySpadeResBlock")
        if self.training and any(hasattr(c, "weight_orig") for c in self._convs()):
            raise NotImplementedError("spectral_norm in training mode runs a power iteration per forward; the native SAMS "
                                      "path is an inference engine -- call .eval()")
        d = {name: ops.PackedConv(effective_weight(c), c.bias, stride=1, pad=c.padding[0], prec=prec)
             for name, c in (("conv_0", self.conv_0), ("conv_1", self.conv_1))}
        if self.learned_shortcut:
            d["conv_s"] = ops.PackedConv(effective_weight(self.conv_s), None, stride=1, pad=0, prec=prec)
        self._packed = (sig, d)
        return d

    def _spade_groups(self):
        """{label key (None for the encoder's single map): [leaf SPADE layers of this block that read it]}."""
        groups = {}
        for mod in [self.spade_0, self.spade_1] + ([self.norm_s] if self.learned_shortcut else []):
            if hasattr(mod, "spade_layers"):
                for key, leaf in mod.spade_layers.items():
                    groups.setdefault(key, []).append(leaf)
            else:
                groups.setdefault(None, []).append(mod)
        return groups

    def _shared_actv(self, ctx, H, W, seg):
        """The mlp_shared convs (spade.py:60-62) of every SPADE layer of the block, one GEMM per label map: all of them read the
        same resized map at the same resolution.  Returns {id(leaf SPADE): its 128-channel window of the activated output}."""
        groups = self._spade_groups()
        if isinstance(seg, torch.Tensor):
            if len(groups) != 1:
                return {}  # MultiSpade.try_fix_labelmap_dict raises the reference's error later
            seg = {next(iter(groups)): seg}
        leaves = [l for ls in groups.values() for l in ls]
        sig = (tuple(params_signature(l.mlp_shared) for l in leaves), ctx.prec)
        if getattr(self, "_packed_shared", None) is None or self._packed_shared[0] != sig:
            d = {key: ops.PackedConv(shared_gemm_weight(ls), torch.cat([l.mlp_shared[0].bias for l in ls], 0), stride=1, pad=0,
                                     prec=ctx.prec) for key, ls in groups.items()}
            self._packed_shared = (sig, d)
        out = {}
        for key, ls in groups.items():
            if key not in seg:
                continue
            a, ap = act_name(ls[0].actvn)
            ks = ls[0].mlp_shared[0].kernel_size[0]
            _, actv = ops.conv2d(ctx.seg_operand(seg[key], H, W, ks), self._packed_shared[1][key], post_act=a, act_param=ap,
                                 want_planes=True)
            for i, leaf in enumerate(ls):
                out[id(leaf)] = actv.window(i * leaf.nhidden, leaf.nhidden)
        return out

    def run(self, ctx, x, seg):
        """spade.py:151-171.  x: f32 NHWC; seg: label map tensor (SPADE) or {name: tensor} (MultiSpade family) -> f32 NHWC."""
        pk = self._pack(ctx.prec)
        act, ap = act_name(self.actvn)
        shared = self._shared_actv(ctx, x.shape[1], x.shape[2], seg)
        if self.learned_shortcut:
            _, ps = self.norm_s.run(ctx, x, seg, shared=shared, want_f32=False, want_planes=True)
            x_s, _ = ops.conv2d(ps, pk["conv_s"], want_f32=True)
        else:
            x_s = x
        _, p0 = self.spade_0.run(ctx, x, seg, act=act, act_param=ap, shared=shared, want_f32=False, want_planes=True)
        dx, _ = ops.conv2d(p0, pk["conv_0"], want_f32=True)
        _, p1 = self.spade_1.run(ctx, dx, seg, act=act, act_param=ap, shared=shared, want_f32=False, want_planes=True)
        dx, _ = ops.conv2d(p1, pk["conv_1"], want_f32=True)
        return ops.add_nhwc(x_s, dx, out=dx)

    precision = None

    def forward(self, x, seg):
        ctx = EngineContext(ops.resolve_precision(self.precision))
        seg = {k: v.contiguous() for k, v in seg.items()} if isinstance(seg, dict) else seg.contiguous()
        return self.run(ctx, x.permute(0, 2, 3, 1).contiguous(), seg).permute(0, 3, 1, 2).contiguous()
