"""WarpModel — Geometric Matching Module (reference: models/warp_model.py:27-152)."""
import argparse
import os
from argparse import ArgumentParser

import torch
from torch import nn

from ..networks.cpvton.warp import (FeatureCorrelation, FeatureExtraction, FeatureL2Norm, FeatureRegression,
                                    TpsGridGen)
from .. import ops
from .base_model import BaseModel, get_and_cat_inputs, maybe_combine_frames_and_channels


# extractionA / extractionB on two streams while a CUDA graph is being captured (parallel graph branches); A/B switch
PARALLEL_EXTRACTION = os.environ.get("SHINEON_GMM_PARALLEL", "0") == "1"


class WarpModel(BaseModel):
    """ Geometric Matching Module """

    @classmethod
    def modify_commandline_options(cls, parser: ArgumentParser, is_train):
        parser = ArgumentParser(parents=[parser], add_help=False)
        parser = super(WarpModel, cls).modify_commandline_options(parser, is_train)
        parser.add_argument("--grid_size", type=int, default=5)
        parser.set_defaults(person_inputs=("agnostic", "cocopose"))
        return parser

    def __init__(self, hparams):
        super().__init__(hparams)
        if isinstance(hparams, dict):
            hparams = argparse.Namespace(**hparams)
        self.extractionA = FeatureExtraction(self.person_channels, ngf=hparams.ngf, n_layers=3,
                                             norm_layer=nn.BatchNorm2d)
        self.extractionB = FeatureExtraction(self.cloth_channels, ngf=hparams.ngf, n_layers=3,
                                             norm_layer=nn.BatchNorm2d)
        self.l2norm = FeatureL2Norm()
        self.correlation = FeatureCorrelation()
        self.regression = FeatureRegression(input_nc=192, output_dim=2 * hparams.grid_size ** 2)
        self.gridGen = TpsGridGen(hparams.fine_height, hparams.fine_width, grid_size=hparams.grid_size)
        self.precision = None

    def set_precision(self, precision):
        """One of ops.PRECISIONS: "fp16x3" (default, fp32-grade), "bf16x3", "fp16", "bf16" (fast modes)."""
        self.precision = precision
        self.extractionA.precision = precision
        self.extractionB.precision = precision
        self.regression.precision = precision

    def regress_theta(self, inputA, inputB):
        """inputA / inputB: f32 NCHW tensors, or the two stems' operands from ops.frame_prep_planes."""
        if PARALLEL_EXTRACTION and torch.cuda.is_current_stream_capturing():
            # the two towers are independent: inside a graph capture they become parallel branches (the capture's private
            # pool is not recycled before the join, so no cross-stream allocator bookkeeping is needed)
            cur = torch.cuda.current_stream()
            side = self.__dict__.get("_side")
            if side is None:
                side = self.__dict__["_side"] = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                featureB = self.extractionB.forward_nhwc(inputB)
            featureA = self.extractionA.forward_nhwc(inputA)
            cur.wait_stream(side)
        else:
            featureA = self.extractionA.forward_nhwc(inputA)
            featureB = self.extractionB.forward_nhwc(inputB)
        _, corr = self.correlation.forward_fused(featureA, featureB, prec=ops.resolve_precision(self.precision))
        return self.regression.forward_planes(corr)

    def forward(self, inputA, inputB):
        """-> (grid [B,H,W,2], theta [B,2*gs^2])  (warp_model.py:63-72)."""
        theta = self.regress_theta(inputA, inputB)
        grid = self.gridGen(theta)
        return grid, theta

    def warp(self, inputA, inputB, cloth, cloth_mask=None, grid_vis=None):
        """forward + the F.grid_sample call sites (warp_model.py:84-86,142-145) with the grid never written
        to memory.  Returns (warped_cloth, warped_mask|None, warped_grid|None, theta)."""
        theta = self.regress_theta(inputA, inputB)
        inputs = [(cloth.contiguous(), "border")]
        if cloth_mask is not None:
            inputs.append((cloth_mask.contiguous(), "zeros"))
        if grid_vis is not None:
            inputs.append((grid_vis.contiguous(), "zeros"))
        outs, _ = self.gridGen.warp(theta, inputs)
        wc = outs[0]
        wm = outs[1] if cloth_mask is not None else None
        wg = outs[-1] if grid_vis is not None else None
        return wc, wm, wg, theta

    def training_step(self, batch, idx, val=False):
        raise NotImplementedError("GMM training (backward kernels) is not part of this build yet (DESIGN.md §9)")

    def test_step(self, batch, batch_idx):
        """Inference step (warp_model.py:115-152) without the PNG writing: returns the warped tensors."""
        batch = maybe_combine_frames_and_channels(self.hparams, batch)
        person_inputs = get_and_cat_inputs(batch, self.hparams.person_inputs)
        cloth_inputs = get_and_cat_inputs(batch, self.hparams.cloth_inputs)
        wc, wm, wg, _ = self.warp(person_inputs, cloth_inputs, batch["cloth"], batch.get("cloth_mask"),
                                  batch.get("grid_vis"))
        self.warped_cloth, self.warped_grid = wc, wg
        return {"warped_cloth": wc, "warped_mask": wm, "warped_grid": wg}
