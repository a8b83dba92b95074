"""UnetMaskModel — CP-VTON Try-On Module (reference: models/unet_mask_model.py:27-326)."""
import argparse
import math
import os

import torch
from torch import nn

from .. import ops
from ..networks import init_weights
from ..networks._engine_util import side_run
from ..networks.cpvton.unet import UnetGenerator
from ..networks.flownet2.native_ops import Resample2d
from ..networks.loss import VGGLoss
from .base_model import BaseModel, get_and_cat_inputs, maybe_combine_frames_and_channels

# VGG features of the target images on the auxiliary stream during the U-Net forward of a training step.  Off: measured on
# B200 at batch 4, two half-batch passes through the slices (targets early, prediction later) cost more than the overlap
# returns -- 5.69 ms against 5.43 ms for the one batched pass over [prediction; target] (profiles/r02_concurrency.md).
PARALLEL_VGG_TARGET = os.environ.get("SHINEON_VGG_TARGET_PARALLEL", "0") == "1"
# the backward's packed operands prepared on the auxiliary stream while the losses are computed: measured 5.57 ms against
# 5.48 ms without (the ~20 pack launches then compete with the loss kernels instead of filling gaps of the backward): off
PREPACK_BACKWARD = os.environ.get("SHINEON_PREPACK_BACKWARD", "0") == "1"


class UnetMaskModel(BaseModel):
    """ CP-VTON Try-On Module (TOM) """

    @classmethod
    def modify_commandline_options(cls, parser: argparse.ArgumentParser, is_train):
        parser = argparse.ArgumentParser(parents=[parser], add_help=False)
        parser = super(UnetMaskModel, cls).modify_commandline_options(parser, is_train)
        parser.set_defaults(person_inputs=("agnostic", "densepose"))
        parser.add_argument("--pen_flow_mask", type=float, default=1.0, help="Penalty applied to flow mask loss")
        return parser

    def __init__(self, hparams):
        super().__init__(hparams)
        if isinstance(hparams, dict):
            hparams = argparse.Namespace(**hparams)
        self.hparams = hparams
        n_frames = hparams.n_frames_total if hasattr(hparams, "n_frames_total") else 1
        self.unet = UnetGenerator(
            input_nc=(self.person_channels + self.cloth_channels) * n_frames,
            output_nc=5 * n_frames if self.hparams.flow_warp else 4 * n_frames,
            num_downs=6,
            num_attention=hparams.num_attn if hasattr(hparams, "num_attn") else 2,
            ngf=int(64 * (math.log(n_frames) + 1)),
            norm_layer=nn.InstanceNorm2d,
            use_self_attn=hparams.self_attn,
            activation=hparams.activation,
        )
        self.resample = Resample2d()
        self.criterionVGG = VGGLoss()  # frozen VGG19 slices (unet_mask_model.py:61); random until a checkpoint is loaded
        init_weights(self.unet, init_type="normal")

    def set_precision(self, precision):
        """One of ops.PRECISIONS: "fp16x3" (default, fp32-grade), "bf16x3", "fp16", "bf16" (fast modes)."""
        self.unet.precision = precision

    def forward(self, person_representation, warped_cloths, flows=None, prev_im=None):
        """-> (p_rendereds, tryon_masks, p_tryons, flow_masks)  (unet_mask_model.py:64-135)."""
        return self._forward(person_representation, warped_cloths, flows)[:4]

    def forward_u8(self, person_representation, warped_cloths, flows=None, f32_outputs=False):
        """The same forward with the try-on frames encoded the way the reference writes them to disk
        (visualization.py:73-76, fused into the compose kernel): returns uint8 [B,n,H,W,3]; with f32_outputs also the
        reference's 4-tuple."""
        res = self._forward(person_representation, warped_cloths, flows, want_u8=True, want_f32=f32_outputs)
        return (res[4],) + tuple(res[:4]) if f32_outputs else res[4]

    def forward_u8_planes(self, unet_in, warped_cloths):
        """forward_u8 when the stem's operand already exists as space-to-depth planes (ops.S2dInput written by
        ops.frame_prep_planes + TpsGridGen.warp_u8): cat([person, warped_cloths], 1) is never materialised."""
        prec = ops.resolve_precision(self.unet.precision)
        assert unet_in.planes.prec == prec, "stem operand precision differs from the model's"
        out = self.unet.run_operand(unet_in, prec)
        return self._compose(out, warped_cloths.contiguous(), None, want_u8=True, want_f32=False)[4]

    def _forward(self, person_representation, warped_cloths, flows=None, want_u8=False, want_f32=True):
        n = self.hparams.n_frames_total
        flow_warp = bool(self.hparams.flow_warp)
        prec = ops.resolve_precision(self.unet.precision)
        person_representation = person_representation.contiguous()
        warped_cloths = warped_cloths.contiguous()
        # torch.cat([person, cloth], 1) is fused into the NCHW -> NHWC-planes (im2col) conversion
        out = self.unet.run_operand((person_representation, warped_cloths), prec)  # f32 NHWC [B,H,W,(4|5)n]
        return self._compose(out, warped_cloths, flows, want_u8, want_f32)

    def _compose(self, out, warped_cloths, flows=None, want_u8=False, want_f32=True):
        n = self.hparams.n_frames_total
        flow_warp = bool(self.hparams.flow_warp)
        B, H, W, _ = out.shape
        dev = out.device
        chain = flows is not None and n > 1  # frame f reads p_tryon of frame f-1 through Resample2d
        p_rendereds = torch.empty(B, 3 * n, H, W, device=dev) if want_f32 else None
        tryon_masks = torch.empty(B, n, H, W, device=dev) if want_f32 else None
        p_tryons = torch.empty(B, 3 * n, H, W, device=dev) if (want_f32 or chain) else None
        flow_masks = torch.empty(B, n, H, W, device=dev) if (flow_warp and want_f32) else None
        tryon_u8 = torch.empty(B, n, H, W, 3, dtype=torch.uint8, device=dev) if want_u8 else None
        outs = (p_rendereds, tryon_masks, p_tryons, flow_masks)
        flows_c = list(torch.chunk(flows, n, dim=1)) if flows is not None else None
        for f in range(n):
            warped_prev = None
            if flows_c is not None and f > 0:
                prev_generated = p_tryons[:, 3 * (f - 1):3 * f].contiguous()
                warped_prev = self.resample(prev_generated, flows_c[f].contiguous())
            ops.tom_compose(out, warped_cloths, n, flow_warp, outs, frame=f, warped_prev=warped_prev, tryon_u8=tryon_u8)
        return p_rendereds, tryon_masks, p_tryons, flow_masks, tryon_u8

    def set_train_precision(self, precision):
        """Numeric mode of the training step: "bf16x3" (default; fp32-grade products, parity-tested) or "bf16" (the
        BASELINE config-5 speed mode: single bf16 tensor-core products, fp32 accumulation / statistics / master weights)."""
        self.unet.precision = precision
        self.criterionVGG.precision = precision

    def training_step(self, batch, batch_idx, val=False):
        """unet_mask_model.py:137-217 without Lightning: forward, the four loss terms and — unless `val` — the whole
        backward pass, all on the hand-written kernels.  Parameter gradients are ACCUMULATED into `.grad` (zero them per
        optimiser step; `accumulated_batches` micro-batches simply call this repeatedly).  Returns
        {"loss": 0-dim tensor, "log": {name: tensor}} with the reference's log names."""
        hp = self.hparams
        if not val:
            self.criterionVGG.warn_if_random()
        batch = maybe_combine_frames_and_channels(hp, batch)
        n = hp.n_frames_total
        flow_warp = bool(hp.flow_warp)
        im, cm = batch["image"].contiguous(), batch["cloth_mask"].contiguous()
        flows = batch["flow"].contiguous() if flow_warp else None
        person = get_and_cat_inputs(batch, hp.person_inputs).contiguous()
        cloths = get_and_cat_inputs(batch, hp.cloth_inputs).contiguous()
        dev = person.device
        # the perceptual loss's target features need nothing from the U-Net: on the auxiliary stream, next to the forward
        # pass (batch 4 leaves most SMs idle in the lower U-Net levels); joined where the loss reads them
        used = [(n - 1, 0.5 if n > 1 else 1.0)] + ([(n - 2, 0.5)] if n > 1 else [])
        tgts = {f: im[:, 3 * f:3 * f + 3].contiguous() for f, _ in used}
        tgt_feats = {}

        def target_branch():
            for f, t in tgts.items():
                tgt_feats[f] = self.criterionVGG.target_features(t)

        join_targets = side_run(target_branch) if PARALLEL_VGG_TARGET else (lambda: None)
        saved_prec = self.unet.precision
        if self.unet.precision is None:
            self.unet.precision = "bf16x3"
        try:
            out = self.unet.forward_train(person, cloths) if not val else self.unet.model.run(
                (person, cloths), ops.resolve_precision(self.unet.precision))
        finally:
            self.unet.precision = saved_prec
        # the backward's packed operands, on the auxiliary stream while the compose / loss kernels run
        join_packs = side_run(self.unet.prepack_backward) if (PREPACK_BACKWARD and not val) else (lambda: None)
        B, H, W, Cout = out.shape
        p_rendereds = torch.empty(B, 3 * n, H, W, device=dev)
        tryon_masks = torch.empty(B, n, H, W, device=dev)
        p_tryons = torch.empty(B, 3 * n, H, W, device=dev)
        flow_masks = torch.empty(B, n, H, W, device=dev) if flow_warp else None
        outs = (p_rendereds, tryon_masks, p_tryons, flow_masks)
        flows_c = [c.contiguous() for c in torch.chunk(flows, n, dim=1)] if flows is not None else None
        prev_gen, warped = [None] * n, [None] * n
        for f in range(n):
            if flows_c is not None and f > 0:
                prev_gen[f] = p_tryons[:, 3 * (f - 1):3 * f].contiguous()
                warped[f] = self.resample(prev_gen[f], flows_c[f])
            ops.tom_compose(out, cloths, n, flow_warp, outs, frame=f, warped_prev=warped[f])
        self.p_rendereds, self.tryon_masks, self.p_tryons, self.flow_masks = outs

        # ---- losses (unet_mask_model.py:173-191): last frame (and the one before it, each weighted 0.5, when n > 1)
        acc = torch.zeros(8, device=dev)  # l1, vgg, mask_l1, flow_mask, then per-frame curr/prev copies for the log
        grad = not val
        join_targets()
        g_tryon = [torch.zeros(B, 3, H, W, device=dev) for _ in range(n)] if grad else [None] * n
        g_mask = [None] * n
        for f, w in used:
            pt = p_tryons[:, 3 * f:3 * f + 3].contiguous()
            tgt = tgts[f]
            ops.l1_loss(pt, tgt, acc[0:1], g_tryon[f], weight=w, accumulate_grad=True)
            self.criterionVGG.loss_and_grad(pt, tgt, acc[1:2], g_tryon[f], scale=w, y_feats=tgt_feats.get(f))
            tm = tryon_masks[:, f:f + 1].contiguous()
            if grad:
                g_mask[f] = torch.empty(B, 1, H, W, device=dev)
            ops.l1_loss(tm, cm[:, f:f + 1].contiguous(), acc[2:3], g_mask[f], weight=w)
        g_fm = None
        if flow_warp:
            last_fm = flow_masks[:, n - 1:n].contiguous()
            ops.channel_sum(last_fm.view(-1, 1), acc[3:4], alpha=float(hp.pen_flow_mask), beta=1.0)
            if grad:
                g_fm = torch.full((B, 1, H, W), float(hp.pen_flow_mask), device=dev)
        loss = acc[0:4].sum()

        # ---- backward: compose (+ flow-warp chain over frames, last to first) -> U-Net
        if grad:
            g_out = torch.zeros(B, H, W, Cout, device=dev)
            for f in reversed(range(n)):
                gw = ops.tom_compose_bwd(out, cloths, n, flow_warp, g_out, frame=f, warped_prev=warped[f],
                                         g_tryons=g_tryon[f], g_masks=g_mask[f],
                                         g_flow_masks=g_fm if f == n - 1 else None, want_g_warped=warped[f] is not None)
                if warped[f] is not None:  # p_tryon of frame f-1 also feeds frame f through Resample2d
                    ops.resample2d_bwd(prev_gen[f], flows_c[f], gw, grad_in1=g_tryon[f - 1])
            join_packs()
            self.unet.backward(g_out)
        if not val:
            self.global_step += 1
        v = "val_" if val else ""
        log = {f"{v}loss/G": loss, f"{v}loss/G/l1": acc[0], f"{v}loss/G/vgg": acc[1], f"{v}loss/G/tryon_mask_l1": acc[2],
               f"{v}loss/G/flow_mask_l1": acc[3]}
        return {"loss": loss, "log": log}

    def test_step(self, batch, batch_idx):
        """Inference step (unet_mask_model.py:250-281) without the PNG writing."""
        batch = maybe_combine_frames_and_channels(self.hparams, batch)
        person_inputs = get_and_cat_inputs(batch, self.hparams.person_inputs)
        cloth_inputs = get_and_cat_inputs(batch, self.hparams.cloth_inputs)
        _, _, self.p_tryon, _ = self.forward(person_inputs, cloth_inputs)
        return {"p_tryon": self.p_tryon[:, -3:]}
