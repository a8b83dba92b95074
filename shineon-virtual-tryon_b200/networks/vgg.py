"""Vgg19 feature slices + VGGLoss (reference: models/networks/vgg.py:6-38, models/networks/loss.py:106-122).

Same module tree as the reference builds out of torchvision's `vgg19().features` (slice1..slice5 with the original
layer indices as child names), so `criterionVGG.vgg.slice3.10.weight` etc. load from reference checkpoints.  The
reference downloads ImageNet weights (`pretrained=True`); there is no network here, so the layers keep their random
initialisation until a checkpoint is loaded.

The perceptual loss is training-only.  Its forward (13 conv3x3+ReLU, 4 max-pools, on the generated and the target
image in one batch), the five weighted L1 terms and the backward to the generated image all run on the hand-written
kernels: tcgen05 conv for every layer and its data gradient, max-pool / ReLU-mask / L1 kernels in between.  The
VGG weights are frozen (vgg.py:30-32), so no weight gradients are formed.
"""
import torch
from torch import nn

from .. import ops
from ._engine_util import params_signature, require_cuda

# torchvision vgg19 "E" configuration up to features[29]
_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512]
_SLICE_ENDS = (2, 7, 12, 21, 30)  # vgg.py:15-24


def _vgg19_features():
    layers, cin = [], 3
    for v in _CFG:
        if v == "M":
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            layers += [nn.Conv2d(cin, v, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
            cin = v
    return layers  # 30 modules: indices 0..29


class Vgg19(nn.Module):
    def __init__(self, requires_grad=False):
        super().__init__()
        feats = _vgg19_features()
        start = 0
        for si, end in enumerate(_SLICE_ENDS, 1):
            seq = nn.Sequential()
            for x in range(start, end):
                seq.add_module(str(x), feats[x])
            setattr(self, f"slice{si}", seq)
            start = end
        if not requires_grad:
            for p in self.parameters():
                p.requires_grad = False
        self._packed = None

    def layers(self):
        """[(kind, module, slice_end?)] in execution order; slice_end marks h_relu1..5."""
        out = []
        for si in range(1, 6):
            mods = list(getattr(self, f"slice{si}").children())
            for j, m in enumerate(mods):
                if isinstance(m, nn.ReLU):
                    continue  # fused into the conv epilogue
                kind = "pool" if isinstance(m, nn.MaxPool2d) else "conv"
                last = all(isinstance(x, nn.ReLU) for x in mods[j + 1:])
                out.append((kind, m, last))
        return out

    def packed(self, prec):
        sig = (params_signature(self), prec)
        if self._packed is None or self._packed[0] != sig:
            require_cuda(self, "Vgg19")
            pk = []
            for kind, m, _ in self.layers():
                if kind != "conv":
                    pk.append(None)
                elif m.in_channels <= 8:
                    pk.append(("i2c", ops.Im2colConv(m.weight, m.bias, 1, 1, prec=prec),
                               ops.PackedConv(m.weight, None, stride=1, pad=1, prec=prec, transposed=True)))
                else:
                    pk.append(("conv", ops.PackedConv(m.weight, m.bias, stride=1, pad=1, prec=prec),
                               ops.PackedConv(m.weight, None, stride=1, pad=1, prec=prec, transposed=True)))
            self._packed = (sig, pk)
        return self._packed[1]

    def forward(self, X):
        """X: f32 NCHW CUDA -> [h_relu1..h_relu5] as f32 NCHW (vgg.py:33-38)."""
        feats, _ = self.run(X.contiguous(), ops.resolve_precision(None), keep=False)
        return [f.permute(0, 3, 1, 2).contiguous() for f in feats]

    def run(self, X, prec, keep):
        """-> ([five f32 NHWC feature maps], tape).  With keep=True the tape holds every layer's f32 output (ReLU
        mask / max-pool routing of the backward)."""
        pk = self.packed(prec)
        feats, tape = [], []
        cur_planes, cur_f32 = None, None
        for (kind, m, last), p in zip(self.layers(), pk):
            if kind == "pool":
                y, yp = ops.maxpool2x2(cur_f32, want_f32=keep, want_planes=True, prec=prec)
                tape.append(("pool", cur_f32, None))
                cur_planes, cur_f32 = yp, y
                continue
            if p[0] == "i2c":
                y, yp = p[1].conv(X, None, post_act="relu", want_f32=True, want_planes=True)
            else:
                y, yp = ops.conv2d(cur_planes, p[1], post_act="relu", want_f32=True, want_planes=True)
            tape.append(("conv", y, p[2]))
            cur_planes, cur_f32 = yp, y
            if last:
                feats.append(y)
        return feats, (tape if keep else None)


class VGGLoss(nn.Module):
    """sum_i w_i * L1(vgg(x)_i, vgg(y)_i.detach()), weights 1/32, 1/16, 1/8, 1/4, 1 (loss.py:106-122)."""

    def __init__(self, layids=None):
        super().__init__()
        self.vgg = Vgg19()
        self.criterion = nn.L1Loss()
        self.weights = [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0]
        self.layids = layids
        self.precision = None
        # The reference builds torchvision.models.vgg19(pretrained=True) (vgg.py:9).  There is no network here, so the
        # slices keep a random initialisation until ImageNet weights arrive through load_pretrained() / a checkpoint
        # that contains criterionVGG.* keys; training on random features is allowed but never silent.
        self.pretrained_loaded = False
        self._warned = False
        self.register_load_state_dict_post_hook(
            lambda module, incompatible: setattr(module, "pretrained_loaded",
                                                 not any("vgg.slice" in k for k in incompatible.missing_keys)))

    def load_pretrained(self, path=None):
        """ImageNet weights for the five slices: `path` = a torchvision vgg19 state_dict file (`features.N.weight` keys,
        e.g. vgg19-dcbff6f7.pth); path=None asks torchvision for its cached download.  Returns True when loaded."""
        if path is None:
            try:
                import torchvision

                sd = torchvision.models.vgg19(weights=torchvision.models.VGG19_Weights.IMAGENET1K_V1).state_dict()
            except Exception:  # noqa: BLE001  (no cache / no network)
                return False
        else:
            sd = torch.load(path, map_location="cpu")
            sd = sd.get("state_dict", sd)
        own = self.vgg.state_dict()
        mapped = {}
        for k in own:  # "slice3.10.weight" <- "features.10.weight"
            src = "features." + k.split(".", 1)[1]
            if src not in sd:
                raise KeyError(f"{src} missing from the VGG19 weights file")
            mapped[k] = sd[src]
        self.vgg.load_state_dict(mapped, strict=True)
        self.pretrained_loaded = True
        return True

    def warn_if_random(self):
        if not self.pretrained_loaded and not self._warned:
            import warnings

            warnings.warn("VGGLoss: the VGG19 slices hold RANDOM weights (the reference uses torchvision's ImageNet weights, "
                          "models/networks/vgg.py:9); the perceptual term is then a random-feature loss.  Pass --vgg_weights "
                          "FILE (torchvision vgg19 state_dict) or load a checkpoint that contains criterionVGG.*", stacklevel=3)
            self._warned = True

    def target_features(self, y):
        """The five feature maps of the (detached) target images, f32 NHWC: they do not depend on the network being trained,
        so a training step may compute them next to its forward pass and hand them to loss_and_grad(y_feats=...)."""
        prec = ops.resolve_precision(self.precision if self.precision is not None else "bf16x3")
        return self.vgg.run(y.contiguous(), prec, keep=False)[0]

    def loss_and_grad(self, x, y, loss, grad_x, scale=1.0, y_feats=None):
        """loss[0] += scale * VGGLoss(x, y); grad_x (f32 NCHW [B,3,H,W], None = value only) += scale * dVGGLoss/dx.
        x, y: f32 NCHW CUDA images.  y_feats = target_features(y) computed earlier (y is then ignored): only x goes
        through the slices here; the values are the same either way (every image is convolved on its own)."""
        prec = ops.resolve_precision(self.precision if self.precision is not None else "bf16x3")
        B = x.shape[0]
        if y_feats is None:
            feats, tape = self.vgg.run(torch.cat([x, y], 0).contiguous(), prec, keep=True)
        else:
            feats, tape = self.vgg.run(x.contiguous(), prec, keep=True)
        ids = list(range(len(feats))) if self.layids is None else list(self.layids)
        # L1 terms: features of x are the first B images of the batch, the targets the last B (contiguous halves)
        g_feat = {}
        for i in ids:
            f = feats[i]
            g = torch.empty_like(f[:B]) if grad_x is not None else None
            ops.l1_loss(f[:B], f[B:] if y_feats is None else y_feats[i], loss, g, weight=scale * self.weights[i], beta_loss=1.0)
            g_feat[id(f)] = g
        if grad_x is None:
            return loss
        # backward through the x half only
        g_cur = None
        for kind, t, dg in reversed(tape):
            if kind == "pool":
                if g_cur is not None:
                    g_cur = ops.maxpool2x2_bwd(t[:B], g_cur)
                continue
            g_here = g_feat.get(id(t))
            if g_cur is None and g_here is None:
                continue  # layers after the last used feature map
            if g_cur is None:
                g1, g2 = g_here, None
            else:
                g1, g2 = g_cur, g_here
            _, G = ops.instnorm_act_bwd(t[:B], None, g1, g2, do_norm=False, act="relu", want_f32=False, want_planes=True,
                                        prec=prec)  # ReLU mask from the layer's own output (y > 0 <=> pre-activation > 0)
            g_cur, _ = ops.conv2d(G, dg, want_f32=True)
        ops.nhwc_to_nchw_add(g_cur, grad_x)
        return loss

    def forward(self, x, y):
        """Loss value only (f32 scalar tensor), for logging / validation."""
        loss = torch.zeros(1, device=x.device)
        self.loss_and_grad(x.contiguous(), y.contiguous(), loss, None)
        return loss[0]
