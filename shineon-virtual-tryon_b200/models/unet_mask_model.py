"""UnetMaskModel — CP-VTON Try-On Module (reference: models/unet_mask_model.py:27-326)."""
import argparse
import math

import torch
from torch import nn

from .. import ops
from ..networks import init_weights
from ..networks.cpvton.unet import UnetGenerator
from ..networks.flownet2.native_ops import Resample2d
from .base_model import BaseModel, get_and_cat_inputs, maybe_combine_frames_and_channels


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
        # VGGLoss (criterionVGG, unet_mask_model.py:61) is training-only and out of this build's scope
        init_weights(self.unet, init_type="normal")

    def set_precision(self, precision):
        """One of ops.PRECISIONS: "fp16x3" (default, fp32-grade), "bf16x3", "fp16", "bf16" (fast modes)."""
        self.unet.precision = precision

    def forward(self, person_representation, warped_cloths, flows=None, prev_im=None):
        """-> (p_rendereds, tryon_masks, p_tryons, flow_masks)  (unet_mask_model.py:64-135)."""
        n = self.hparams.n_frames_total
        flow_warp = bool(self.hparams.flow_warp)
        prec = ops.resolve_precision(self.unet.precision)
        person_representation = person_representation.contiguous()
        warped_cloths = warped_cloths.contiguous()
        # torch.cat([person, cloth], 1) is fused into the NCHW -> NHWC-planes (im2col) conversion
        out = self.unet.model.run((person_representation, warped_cloths), prec)  # f32 NHWC [B,H,W,(4|5)n]
        B, H, W, _ = out.shape
        dev = out.device
        p_rendereds = torch.empty(B, 3 * n, H, W, device=dev)
        tryon_masks = torch.empty(B, n, H, W, device=dev)
        p_tryons = torch.empty(B, 3 * n, H, W, device=dev)
        flow_masks = torch.empty(B, n, H, W, device=dev) if flow_warp else None
        outs = (p_rendereds, tryon_masks, p_tryons, flow_masks)
        flows_c = list(torch.chunk(flows, n, dim=1)) if flows is not None else None
        for f in range(n):
            warped_prev = None
            if flows_c is not None and f > 0:
                prev_generated = p_tryons[:, 3 * (f - 1):3 * f].contiguous()
                warped_prev = self.resample(prev_generated, flows_c[f].contiguous())
            ops.tom_compose(out, warped_cloths, n, flow_warp, outs, frame=f, warped_prev=warped_prev)
        return p_rendereds, tryon_masks, p_tryons, flow_masks

    def training_step(self, batch, batch_idx, val=False):
        raise NotImplementedError("U-Net training (backward kernels, VGG loss) is not part of this build yet "
                                  "(DESIGN.md §9)")

    def test_step(self, batch, batch_idx):
        """Inference step (unet_mask_model.py:250-281) without the PNG writing."""
        batch = maybe_combine_frames_and_channels(self.hparams, batch)
        person_inputs = get_and_cat_inputs(batch, self.hparams.person_inputs)
        cloth_inputs = get_and_cat_inputs(batch, self.hparams.cloth_inputs)
        _, _, self.p_tryon, _ = self.forward(person_inputs, cloth_inputs)
        return {"p_tryon": self.p_tryon[:, -3:]}
