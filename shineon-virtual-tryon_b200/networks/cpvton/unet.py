"""UnetGenerator / UnetSkipConnectionBlock (reference: models/networks/cpvton/unet.py:9-211).

The module tree (and therefore every state_dict key, e.g. `model.model.1.model.3.model.1.weight`) is built
exactly like the reference constructors do.  The forward pass does not run the nn.Sequential: it walks the
tree once and drives the hand-written kernels, keeping activations as NHWC bf16 hi/lo planes between layers:

  down:  conv4x4s2 (tcgen05, bias in epilogue) -> InstanceNorm(+next activation) -> [SelfAttention]
  up:    upsample2x_cat (up_act, concat, bilinear x2) -> conv3x3 (tcgen05) -> InstanceNorm -> [SelfAttention]
"""
import torch
from torch import nn

from ... import ops
from .._engine_util import side_run, fold_bn, params_signature, require_cuda
from ..activation import Sine, Swish, act_name
from ..attention.sagan import SelfAttention


class UnetGenerator(nn.Module):
    def __init__(self, input_nc, output_nc, num_downs, num_attention, ngf=64, norm_layer=nn.BatchNorm2d,
                 use_dropout=False, use_self_attn=False, activation=None):
        super().__init__()

        def attn():
            return use_self_attn if use_self_attn and num_attention > 0 else None

        unet_block = UnetSkipConnectionBlock(ngf * 8, ngf * 8, input_nc=None, submodule=None, norm_layer=norm_layer,
                                             innermost=True, self_attn=attn(), activation=activation)
        num_attention -= 1
        for _ in range(num_downs - 5):
            unet_block = UnetSkipConnectionBlock(ngf * 8, ngf * 8, input_nc=None, submodule=unet_block,
                                                 norm_layer=norm_layer, use_dropout=use_dropout, self_attn=attn(),
                                                 activation=activation)
            num_attention -= 1
        for outer, inner in ((ngf * 4, ngf * 8), (ngf * 2, ngf * 4), (ngf, ngf * 2)):
            unet_block = UnetSkipConnectionBlock(outer, inner, input_nc=None, submodule=unet_block,
                                                 norm_layer=norm_layer, self_attn=attn(), activation=activation)
            num_attention -= 1
        unet_block = UnetSkipConnectionBlock(output_nc, ngf, input_nc=input_nc, submodule=unet_block, outermost=True,
                                             norm_layer=norm_layer, self_attn=attn(), activation=activation)
        self.model = unet_block
        self.precision = None  # ops.PRECISIONS name; None = default fp16x3 (fp32-grade products)

    def forward_nhwc(self, input):
        """input: f32 NCHW CUDA tensor -> f32 NHWC [N,H,W,output_nc]."""
        require_cuda(self, "UnetGenerator")
        prec = ops.resolve_precision(self.precision)
        return self.run_operand((input.contiguous(), None), prec)

    def run_operand(self, a_in, prec):
        """The inference pass on (x0, x1|None) f32 NCHW inputs or a prebuilt stem operand (ops.S2dInput)."""
        N = a_in.N if isinstance(a_in, ops.S2dInput) else a_in[0].shape[0]
        dev = a_in.planes.hi.device if isinstance(a_in, ops.S2dInput) else a_in[0].device
        blk = UnetSkipConnectionBlock
        blk._ARENA = [torch.zeros(self.model.stats_arena_size(N), dtype=torch.float64, device=dev), 0] if blk.FUSE_STATS else None
        try:
            return self.model.run(a_in, prec)
        finally:
            blk._ARENA = None

    def forward(self, input):
        return self.forward_nhwc(input).permute(0, 3, 1, 2).contiguous()

    # ------------------------------------------------------------------ training
    def forward_train(self, x0, x1=None):
        """Forward that keeps what the hand-written backward needs (conv inputs as planes, pre-norm conv outputs,
        InstanceNorm statistics, attention projections).  (x0, x1): f32 NCHW inputs, concatenated on channels.
        Returns the f32 NHWC output; call `backward(grad_nhwc)` next."""
        require_cuda(self, "UnetGenerator")
        prec = ops.resolve_precision(self.precision if self.precision is not None else "bf16x3")
        self._train_prec = prec
        return self.model.run((x0.contiguous(), None if x1 is None else x1.contiguous()), prec, train=True)

    def prepack_backward(self):
        """Packs the backward pass's conv operands (flipped / transposed 16-bit weights of every block and attention layer)
        ahead of time; a training step calls it on the auxiliary stream while the losses are being computed, so the first
        backward after an optimiser step does not stop for ~20 small pack launches.  backward() finds them cached."""
        prec = self._train_prec
        blk = self.model
        while blk is not None:
            blk._pack_bwd(prec)
            for a in (blk._parts["attn_up"], blk._parts["attn_down"]):
                if a is not None:
                    a.packed_dgrad(prec)
            blk = blk._parts["sub"]

    def backward(self, grad_out_nhwc):
        """Accumulates dL/dparam into every parameter's .grad (allocated on first use) given dL/d(output) as f32 NHWC
        [N,H,W,output_nc].  The input gradient is not produced (the U-Net's inputs are data)."""
        self.model.backward(grad_out_nhwc.contiguous(), None, self._train_prec)


# Set by training.Trainer for the micro-batch whose gradients get exchanged: called with the parameters whose
# gradients just became final, so the data-parallel all-reduce of a full bucket overlaps the rest of the backward.
GRAD_READY_HOOK = None


def _ready(params):
    if GRAD_READY_HOOK is not None:
        GRAD_READY_HOOK([p for p in params if p is not None])


def _grad_of(p):
    """The parameter's gradient accumulator (zero-initialised on first use; a view of the flat data-parallel
    gradient buffer when ops-level training helpers installed one)."""
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return p.grad


class UnetSkipConnectionBlock(nn.Module):
    def __init__(self, outer_nc, inner_nc, input_nc=None, submodule=None, outermost=False, innermost=False,
                 norm_layer=nn.BatchNorm2d, self_attn=False, use_dropout=False, activation=None):
        super().__init__()
        self.outermost = outermost
        self.innermost = innermost
        use_bias = norm_layer == nn.InstanceNorm2d
        if input_nc is None:
            input_nc = outer_nc
        downconv = nn.Conv2d(input_nc, inner_nc, kernel_size=4, stride=2, padding=1, bias=use_bias)
        down_activation = nn.LeakyReLU(0.2, True) if activation is None else _get_activation_fn(activation)
        downnorm = norm_layer(inner_nc)
        up_activation = nn.ReLU(True) if activation is None else _get_activation_fn(activation)
        upnorm = norm_layer(outer_nc)
        upsample = nn.Upsample(scale_factor=2, mode="bilinear")
        if outermost:
            upconv = nn.Conv2d(inner_nc * 2, outer_nc, kernel_size=3, stride=1, padding=1, bias=use_bias)
            down = [downconv]
            up = [up_activation, upsample, upconv, upnorm]
            if self_attn:
                down.append(SelfAttention(inner_nc, "relu"))
                up.append(SelfAttention(outer_nc, "relu"))
            model = down + [submodule] + up
        elif innermost:
            upconv = nn.Conv2d(inner_nc, outer_nc, kernel_size=3, stride=1, padding=1, bias=use_bias)
            down = [down_activation, downconv]
            up = [up_activation, upsample, upconv, upnorm]
            if self_attn:
                down.append(SelfAttention(inner_nc, "relu"))
                up.append(SelfAttention(outer_nc, "relu"))
            model = down + up
        else:
            upconv = nn.Conv2d(inner_nc * 2, outer_nc, kernel_size=3, stride=1, padding=1, bias=use_bias)
            down = [down_activation, downconv, downnorm]
            up = [up_activation, upsample, upconv, upnorm]
            if self_attn:
                down.append(SelfAttention(inner_nc, "relu"))
                up.append(SelfAttention(outer_nc, "relu"))
            model = down + [submodule] + up + ([nn.Dropout(0.5)] if use_dropout else [])
        self.model = nn.Sequential(*model)
        # structural view used by the engine (not registered twice: plain attributes to existing children)
        self._parts = dict(
            down_act=None if outermost else down_activation, downconv=downconv,
            downnorm=None if (outermost or innermost) else downnorm,
            attn_down=down[-1] if self_attn else None, sub=submodule, up_act=up_activation, upconv=upconv,
            upnorm=upnorm, attn_up=up[-1] if self_attn else None, default_act=activation is None,
            dropout=use_dropout and not (outermost or innermost))
        self._packed = None

    # ------------------------------------------------------------------ engine
    def _pack(self, prec):
        sig = (params_signature(self._parts["downconv"]) + params_signature(self._parts["upconv"])
               + params_signature(self._parts["upnorm"])
               + (params_signature(self._parts["downnorm"]) if self._parts["downnorm"] is not None else ()), prec)
        if self._packed is not None and self._packed[0] == sig:
            return self._packed[1]
        pr = self._parts
        dc, uc = pr["downconv"], pr["upconv"]
        sub = pr["sub"]
        d = {}
        d["down_i2c"] = None
        if self.outermost and dc.in_channels <= 32:
            # tiny Cin: im2col'd input + dense 1x1 GEMM instead of one mostly-zero K-block per tap
            d["down_i2c"] = ops.Im2colConv(dc.weight, dc.bias, 2, 1, prec=prec)  # training keeps the im2col operand
            d["down"] = d["down_i2c"].pc
            d["down_first"] = ops.first_layer_conv(dc.weight, dc.bias, 2, 1, prec=prec)
        else:
            d["down"] = ops.PackedConv(dc.weight, dc.bias, stride=2, pad=1, prec=prec)
        if sub is not None:
            # up-conv input channels: [skip (sub input, padded to 64) | x' (sub output, padded to 64)]
            c_skip = sub._parts["downconv"].in_channels
            c_xp = sub._parts["upconv"].out_channels
            p_skip, p_xp = ops.cpad64(c_skip), ops.cpad64(c_xp)
            # device-resident once per block: re-packing after every optimiser step must not copy from the host
            # (the training step is CUDA-graph captured)
            dev = uc.weight.device
            if getattr(self, "_cmap_dev", None) is None or self._cmap_dev.device != dev:
                cm = [-1] * (p_skip + p_xp)
                for c in range(c_skip):
                    cm[c] = c
                for c in range(c_xp):
                    cm[p_skip + c] = c_skip + c
                self._cmap_dev = torch.as_tensor(cm, dtype=torch.int32, device=dev)
            cmap = self._cmap_dev
            d["up_tap"] = None
            if self.outermost and uc.out_channels <= 8 and isinstance(pr["upnorm"], nn.InstanceNorm2d):
                # few output channels: tap-stacked 1x1 GEMM + col2im (the activation is read once, not once per tap)
                d["up_tap"] = ops.TapStackedConv3x3(uc.weight, uc.bias, prec=prec, cin_pad=p_skip + p_xp, chan_map=cmap)
            d["up"] = ops.PackedConv(uc.weight, uc.bias, stride=1, pad=1, prec=prec, cin_pad=p_skip + p_xp,
                                     chan_map=cmap)
            d["cat_geom"] = (c_skip, p_skip, c_xp, p_xp)
        else:
            d["up_tap"] = None
            d["up"] = ops.PackedConv(uc.weight, uc.bias, stride=1, pad=1, prec=prec)
        # inference: upsample -> 3x3 conv evaluated at the low resolution (ops.UpsampledConv3x3).  Not for the
        # default-activation quirk (skip and child input differ by a ReLU there) nor folded BatchNorm epilogues.
        d["up_lowres"] = None  # packed on first inference use (_lowres); the training tape never needs it
        d["lowres_ok"] = not pr["default_act"] and isinstance(pr["upnorm"], nn.InstanceNorm2d)
        d["cmap_up"] = None
        if sub is not None:
            d["cmap_up"] = cmap
        d["up_dgrad"] = d["down_dgrad"] = None  # packed lazily by _pack_bwd (training only)
        for key, norm in (("down_bn", pr["downnorm"]), ("up_bn", pr["upnorm"])):
            d[key] = None
            if isinstance(norm, nn.BatchNorm2d):
                d[key] = fold_bn(norm)
            elif norm is not None and not isinstance(norm, nn.InstanceNorm2d):
                raise NotImplementedError(f"norm layer {type(norm).__name__} has no native kernel")
            elif isinstance(norm, nn.InstanceNorm2d) and (norm.affine or norm.track_running_stats):
                raise NotImplementedError("InstanceNorm2d with affine/running stats is not used by the reference")
        self._packed = (sig, d)
        return d

    def _lowres(self, pk, prec):
        if not (ops.UPCONV_LOWRES and pk["lowres_ok"]):
            return None
        if pk["up_lowres"] is None:
            uc = self._parts["upconv"]
            if self._parts["sub"] is not None:
                _, p_skip, _, p_xp = pk["cat_geom"]
                pk["up_lowres"] = ops.UpsampledConv3x3(uc.weight, uc.bias, prec=prec, cin_pad=p_skip + p_xp,
                                                       chan_map=pk["cmap_up"])
            else:
                pk["up_lowres"] = ops.UpsampledConv3x3(uc.weight, uc.bias, prec=prec)
        return pk["up_lowres"]

    def _pack_bwd(self, prec):
        """Data-gradient operands: a stride-1 conv's dgrad is the flipped-tap conv with Cin/Cout swapped, the 4x4 s2
        down-conv's dgrad is the four-phase transposed conv (both on the same tcgen05 kernel as the forward)."""
        from ..deconv import PackedDeconv4x4s2

        pk = self._pack(prec)
        if pk["up_dgrad"] is None:
            pr = self._parts
            pk["up_dgrad"] = ops.PackedConv(pr["upconv"].weight, None, stride=1, pad=1, prec=prec, transposed=True)
            if not self.outermost:
                pk["down_dgrad"] = PackedDeconv4x4s2(pr["downconv"].weight, None, prec=prec)
        return pk

    def _finish_train(self, conv_f32, norm, attn, act, act_param, prec, want_final_f32=False):
        """Training form of _finish: nothing is overwritten in place and the backward's inputs are returned."""
        inorm = isinstance(norm, nn.InstanceNorm2d)
        eps = norm.eps if inorm else 1e-5
        ws = ops.instnorm_stats_ws(conv_f32) if inorm else None
        sv = dict(c=conv_f32, ws=ws, inorm=inorm, eps=eps, act=act, act_param=act_param, attn=attn)
        if attn is None:
            if want_final_f32:
                y, _ = ops.instnorm_act(conv_f32, do_norm=inorm, eps=eps, act=None, want_f32=True, want_planes=False, ws=ws)
                sv["act"] = None
                return y, sv
            _, p = ops.instnorm_act(conv_f32, do_norm=inorm, eps=eps, act=act, act_param=act_param, want_f32=False,
                                    want_planes=True, prec=prec, ws=ws)
            return p, sv
        y, p = ops.instnorm_act(conv_f32, do_norm=inorm, eps=eps, act=None, want_f32=True, want_planes=True, prec=prec, ws=ws)
        z, qkv = attn.run_train(y, p)
        sv.update(p=p, qkv=qkv, z=z)
        if want_final_f32:
            sv["act"] = None
            return z, sv
        _, out = ops.instnorm_act(z, do_norm=False, act=act, act_param=act_param, want_f32=False, want_planes=True, prec=prec)
        return out, sv

    @staticmethod
    def _finish_bwd(sv, g1, g2, prec):
        """-> (f32 NHWC, Planes) of dL/d(conv output) given dL/d(activated output) = g1 (+ g2)."""
        attn = sv["attn"]
        if attn is None:
            return ops.instnorm_act_bwd(sv["c"], sv["ws"], g1, g2, do_norm=sv["inorm"], act=sv["act"],
                                        act_param=sv["act_param"], eps=sv["eps"], want_f32=True, want_planes=True, prec=prec)
        gz = g1 if (sv["act"] is None and g2 is None) else ops.act_bwd(sv["z"], g1, g2, act=sv["act"], act_param=sv["act_param"])
        g_y = attn.backward(sv["p"], sv["qkv"], gz, prec)  # through the q|k|v projections; the residual is gz itself
        return ops.instnorm_act_bwd(sv["c"], sv["ws"], gz, g_y, do_norm=sv["inorm"], act=None, eps=sv["eps"],
                                    want_f32=True, want_planes=True, prec=prec)

    def _finish(self, conv_f32, norm, bn, attn, act, act_param, prec, want_final_f32=False, out_planes=None, ws=None):
        """conv output (f32 NHWC, bias [and folded BN] applied) -> [InstanceNorm] -> [SelfAttention] -> act.
        Returns Planes of the activated value, or the final f32 tensor when want_final_f32.
        ws: InstanceNorm statistics already accumulated by the kernel that produced conv_f32 (_stats_ws)."""
        inorm = isinstance(norm, nn.InstanceNorm2d)
        st = dict(ws=ws, stats_ready=True) if (inorm and ws is not None) else {}
        if attn is None:
            if want_final_f32:
                y, _ = ops.instnorm_act(conv_f32, do_norm=inorm, eps=norm.eps if inorm else 1e-5, act=None,
                                        want_f32=True, want_planes=False, out_f32=conv_f32, **st)
                return y
            _, p = ops.instnorm_act(conv_f32, do_norm=inorm, eps=norm.eps if inorm else 1e-5, act=act,
                                    act_param=act_param, want_f32=False, want_planes=True, prec=prec,
                                    out_planes=out_planes, **st)
            return p
        # attention works on the normalised, un-activated tensor: needs it as f32 (residual) and planes (qkv conv)
        y, p = ops.instnorm_act(conv_f32, do_norm=inorm, eps=norm.eps if inorm else 1e-5, act=None, want_f32=True,
                                want_planes=True, prec=prec, out_f32=conv_f32, **st)
        if want_final_f32:
            return attn.run(y, p, want_f32=True, want_planes=False)[0]
        return attn.run(y, p, act=act, act_param=act_param, want_f32=False, want_planes=True, out_planes=out_planes)[1]

    # InstanceNorm statistics are accumulated by the kernel that PRODUCES the tensor (conv epilogue / upconv gather):
    # False = the separate statistics pass of instnorm_act (kept for A/B measurements and the tests)
    FUSE_STATS = True

    # one zero-filled f64 arena per forward pass (UnetGenerator.forward_nhwc sets it up): the blocks carve their
    # statistics buffers out of it instead of launching one fill kernel per normalisation
    _ARENA = None

    @classmethod
    def _stats_ws(cls, norm, N, C, device):
        if not (cls.FUSE_STATS and isinstance(norm, nn.InstanceNorm2d)):
            return None
        n = 2 * N * C
        ar = cls._ARENA
        if ar is not None and ar[0].device == device and ar[1] + n <= ar[0].numel():
            ws = ar[0][ar[1]:ar[1] + n]
            ar[1] += n
            return ws
        return torch.zeros(n, dtype=torch.float64, device=device)

    def stats_arena_size(self, N):
        """f64 elements the statistics buffers of one pass through this block (and its children) need for batch N."""
        pr = self._parts
        n = 0
        for norm, conv in ((pr["downnorm"], pr["downconv"]), (pr["upnorm"], pr["upconv"])):
            if isinstance(norm, nn.InstanceNorm2d):
                n += 2 * N * conv.out_channels
        return n + (pr["sub"].stats_arena_size(N) if pr["sub"] is not None else 0)

    def run(self, a_in, prec, train=False, out=None, raw_out=False):
        """a_in: Planes holding this block's (already down-activated) input; for the outermost block a tuple
        (x0, x1|None) of f32 NCHW tensors (torch.cat([x0, x1], 1) is fused into the layout conversion).
        Outermost (or raw_out): returns the f32 NHWC output x'.  Otherwise returns Planes of up_act(x') for the parent."""
        pr = self._parts
        if self.training and pr["dropout"]:
            raise NotImplementedError("Dropout in training mode is not implemented in the native U-Net engine")
        pk = self._pack(prec)
        sub = pr["sub"]
        if train:
            return self._run_train(a_in, prec, pk)
        up_act, up_par = act_name(pr["up_act"])
        # ---- down path: conv -> [norm] -> [attn] -> activation consumed next
        if self.innermost:
            next_act, next_par = up_act, up_par
        else:
            next_act, next_par = act_name(sub._parts["down_act"])
        bn = pk["down_bn"]
        sc, sh = bn if bn is not None else (None, None)
        dc, uc = pr["downconv"], pr["upconv"]
        low = self._lowres(pk, prec)
        s2d_in = a_in if isinstance(a_in, ops.S2dInput) else None  # the stem's operand, built by ops.frame_prep_planes
        if s2d_in is not None:
            assert self.outermost and isinstance(pk.get("down_first"), ops.S2dConv) and s2d_in.C == pk["down_first"].Cin, \
                "S2dInput does not match this U-Net's first layer"
            N_, dev, hin, win = s2d_in.N, s2d_in.planes.hi.device, s2d_in.H, s2d_in.W
        elif isinstance(a_in, tuple):
            N_, dev = a_in[0].shape[0], a_in[0].device
            hin, win = a_in[0].shape[2:]
        else:
            N_, dev, hin, win = a_in.N, a_in.hi.device, a_in.H, a_in.W
        # a down conv with neither norm nor attention behind it (outermost / innermost block): bias + the next layer's
        # activation in the conv epilogue, planes written straight into the concat buffer -- no f32 round trip
        direct_planes = (low is not None and pr["downnorm"] is None and pr["attn_down"] is None and bn is None
                         and (self.outermost or self.innermost))
        cat = None
        if low is not None and not self.innermost:
            c_skip, p_skip, c_xp, p_xp = pk["cat_geom"]
            cat = ops.Planes(N_, hin // 2, win // 2, c_skip + c_xp, prec=prec, device=dev, cpad=p_skip + p_xp)
        ws_d = self._stats_ws(pr["downnorm"], N_, dc.out_channels, dev)
        if direct_planes:
            okw = dict(post_act=next_act, act_param=next_par,
                       **(dict(out_planes=cat.window(0, pk["cat_geom"][0])) if cat is not None else dict(want_planes=True)))
        else:
            okw = dict(scale=sc, shift=sh, want_f32=True)
        if s2d_in is not None:
            f32, a_mid = pk["down_first"].conv_planes(s2d_in.planes, **okw)
        elif isinstance(a_in, tuple) and pk["down_i2c"] is not None:
            f32, a_mid = pk["down_first"].conv(a_in[0], a_in[1], **okw)
        else:
            if isinstance(a_in, tuple):
                a_in = ops.nchw_to_planes(a_in[0], a_in[1], prec=prec)
            f32, a_mid = ops.conv2d(a_in, pk["down"], stats_ws=ws_d, **okw)
        if low is not None:
            # up-path operand at the LOW resolution: [skip | x'] written straight into one concat buffer by their producers
            if self.innermost:
                cat = a_mid if direct_planes else self._finish(f32, pr["downnorm"], bn, pr["attn_down"], next_act, next_par,
                                                               prec, ws=ws_d)
            else:
                if not direct_planes:
                    a_mid = self._finish(f32, pr["downnorm"], bn, pr["attn_down"], next_act, next_par, prec,
                                         out_planes=cat.window(0, c_skip), ws=ws_d)
                sub.run(a_mid, prec, out=cat.window(p_skip, c_xp))
            ws_u = self._stats_ws(pr["upnorm"], N_, uc.out_channels, dev)
            f32 = low(cat, stats_ws=ws_u)
            if self.outermost or raw_out:
                return self._finish(f32, pr["upnorm"], bn, pr["attn_up"], None, 0.0, prec, want_final_f32=True, ws=ws_u)
            return self._finish(f32, pr["upnorm"], bn, pr["attn_up"], up_act, up_par, prec, out_planes=out, ws=ws_u)
        a_mid = self._finish(f32, pr["downnorm"], bn, pr["attn_down"], next_act, next_par, prec, ws=ws_d)
        # ---- child + up-path input
        if self.innermost:
            u = ops.upsample2x_cat(a_mid, None)
        else:
            xp = sub.run(a_mid, prec)
            # default activation: the skip holds LeakyReLU'd values (in-place, unet.py:132) and the parent's
            # ReLU is applied on top when reading it (relu(leaky(x)) == relu(x)); x' is already ReLU'd (idempotent)
            extra = "relu" if pr["default_act"] else None
            u = ops.upsample2x_cat(a_mid, xp, act=extra)
        bn = pk["up_bn"]
        sc, sh = bn if bn is not None else (None, None)
        if pk["up_tap"] is not None:
            f32 = pk["up_tap"](u)
        else:
            f32, _ = ops.conv2d(u, pk["up"], scale=sc, shift=sh, want_f32=True)
        if self.outermost or raw_out:
            return self._finish(f32, pr["upnorm"], bn, pr["attn_up"], None, 0.0, prec, want_final_f32=True)
        return self._finish(f32, pr["upnorm"], bn, pr["attn_up"], up_act, up_par, prec, out_planes=out)

    # ------------------------------------------------------------------ training engine (row U6)
    def _run_train(self, a_in, prec, pk):
        pr = self._parts
        sub = pr["sub"]
        if pr["default_act"]:
            raise NotImplementedError("training the default-activation U-Net (in-place LeakyReLU skip quirk, unet.py:132) "
                                      "has no native backward; the ShineOn recipe trains with --activation gelu")
        if pk["down_bn"] is not None or pk["up_bn"] is not None:
            raise NotImplementedError("BatchNorm2d U-Net training (batch statistics) has no native kernel; the reference's "
                                      "UnetMaskModel uses InstanceNorm2d (unet_mask_model.py:56)")
        up_act, up_par = act_name(pr["up_act"])
        next_act, next_par = (up_act, up_par) if self.innermost else act_name(sub._parts["down_act"])
        if isinstance(a_in, tuple):
            if pk["down_i2c"] is not None:
                a_in = pk["down_i2c"].prepare(a_in[0], a_in[1])  # im2col'd planes: the GEMM operand, kept for wgrad
            else:
                a_in = ops.nchw_to_planes(a_in[0], a_in[1], prec=prec)
        f32, _ = ops.conv2d(a_in, pk["down"], want_f32=True)
        a_mid, sv_down = self._finish_train(f32, pr["downnorm"], pr["attn_down"], next_act, next_par, prec)
        xp = None if self.innermost else sub.run(a_mid, prec, train=True)
        u = ops.upsample2x_cat(a_mid, xp)
        f32, _ = ops.conv2d(u, pk["up"], want_f32=True)
        out, sv_up = self._finish_train(f32, pr["upnorm"], pr["attn_up"], up_act, up_par, prec,
                                        want_final_f32=self.outermost)
        self._tape = dict(a_in=a_in, down=sv_down, u=u, up=sv_up)
        return out

    def up_params(self):
        pr = self._parts
        ps = [pr["upconv"].weight, pr["upconv"].bias]
        if pr["attn_up"] is not None:
            ps += list(pr["attn_up"].parameters())
        return [p for p in ps if p is not None]

    def down_params(self):
        pr = self._parts
        ps = [pr["downconv"].weight, pr["downconv"].bias]
        if pr["attn_down"] is not None:
            ps += list(pr["attn_down"].parameters())
        return [p for p in ps if p is not None]

    def backward(self, g1, g2, prec):
        """g1 (+ g2): dL/d(this block's returned value), f32 NHWC.  Accumulates parameter gradients; returns
        dL/d(block input planes) as f32 NHWC (None for the outermost block)."""
        pr, tp = self._parts, self._tape
        self._tape = None
        pk = self._pack_bwd(prec)
        sub = pr["sub"]
        dc, uc = pr["downconv"], pr["upconv"]
        # ---- up path: norm/attn/act -> conv3x3
        # (weight / bias gradients on the auxiliary stream, see _engine_util.side_run; gc_* / G_* / tp stay referenced until
        # the joins at the end.  With a gradient-ready hook installed -- eager data-parallel overlap -- everything stays on
        # one stream so that "ready" keeps meaning "enqueued before this point".)
        par = GRAD_READY_HOOK is None
        gc_u, G_u = self._finish_bwd(tp["up"], g1, g2, prec)

        def up_wgrad():
            if uc.bias is not None:
                ops.channel_sum(gc_u, _grad_of(uc.bias), beta=1.0)
            ops.conv2d_wgrad(G_u, tp["u"], _grad_of(uc.weight), Cout=uc.out_channels, Cin=uc.in_channels, kh=3, kw=3,
                             stride=1, pad=1, chan_map=pk["cmap_up"], beta=1.0)

        join_u = side_run(up_wgrad, par)
        _ready(self.up_params())
        g_u, _ = ops.conv2d(G_u, pk["up_dgrad"], want_f32=True)  # [N,2H,2W,Cin_up]
        # ---- bilinear x2 + concat
        if sub is None:
            g_skip, _ = ops.upsample2x_cat_bwd(g_u, uc.in_channels, 0)
            g_child = None
        else:
            c_skip = sub._parts["downconv"].in_channels
            g_skip, g_xp = ops.upsample2x_cat_bwd(g_u, c_skip, uc.in_channels - c_skip)
            g_child = sub.backward(g_xp, None, prec)
        # ---- down path: norm/attn/act -> conv4x4 s2
        gc_d, G_d = self._finish_bwd(tp["down"], g_skip, g_child, prec)

        def down_wgrad():
            if dc.bias is not None:
                ops.channel_sum(gc_d, _grad_of(dc.bias), beta=1.0)
            ops.conv2d_wgrad(G_d, tp["a_in"], _grad_of(dc.weight), Cout=dc.out_channels, Cin=dc.in_channels, kh=4, kw=4,
                             stride=2, pad=1, mode=1 if pk["down_i2c"] is not None else 0, beta=1.0)

        join_d = side_run(down_wgrad, par)
        _ready(self.down_params())
        g_in = None if self.outermost else pk["down_dgrad"](G_d, want_f32=True)[0]
        join_u()
        join_d()
        return g_in

    precision = None  # ops.PRECISIONS name for a block used on its own (UnetGenerator passes its own)

    def forward(self, x):
        """unet.py:188-198 for a block used on its own (inference): outermost -> model(x); otherwise
        torch.cat([x, model(x)], 1).  x: f32 NCHW CUDA.  With the default activation the reference's in-place
        LeakyReLU(0.2, True) (unet.py:132) also rewrites the caller's x, so the skip half holds leaky(x): reproduced."""
        require_cuda(self, "UnetSkipConnectionBlock")
        if self.training and self._parts["dropout"]:
            raise NotImplementedError("Dropout in training mode is not implemented in the native U-Net engine")
        prec = ops.resolve_precision(self.precision)
        x = x.contiguous()
        if self.outermost:
            return self.run((x, None), prec).permute(0, 3, 1, 2).contiguous()
        act, par = act_name(self._parts["down_act"])
        a_in = ops.nchw_to_planes(x, act=act, act_param=par, prec=prec)
        xp = self.run(a_in, prec, raw_out=True).permute(0, 3, 1, 2)
        if self._parts["default_act"]:
            flat, _ = ops.instnorm_act(x.view(1, 1, -1, 8) if x.numel() % 8 == 0 else x.view(1, 1, -1, 1), do_norm=False,
                                       act=act, act_param=par, want_f32=True, want_planes=False)
            x.copy_(flat.view_as(x))  # the in-place side effect on the caller's tensor
        return torch.cat([x, xp], 1)


def _get_activation_fn(activation):
    if activation == "relu":
        return nn.ReLU()
    elif activation == "gelu":
        return nn.GELU()
    elif activation == "swish":
        return Swish()
    elif activation == "sine":
        return Sine()
    else:
        raise RuntimeError(f"The selected activation should be relu/gelu/swish/sine, not {activation}")
