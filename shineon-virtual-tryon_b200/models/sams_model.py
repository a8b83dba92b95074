"""SamsModel — Self-Attentive Multi-SPADE model, generator side (reference: models/sams_model.py:23-271).

What the hot path needs of it: the option surface, the generator, and `generate_n_frames` — the autoregressive frame loop
(previous-frame window -> SamsGenerator -> optional Resample2d flow blend).  The GAN training losses and discriminators
are out of scope; the reference has no test-time path for this model either (test_step is `pass`,
sams_model.py:346-347), so `generate_n_frames` is the inference entry point here too.
"""
import argparse

import torch

from .. import ops
from ..networks.flownet2.native_ops import Resample2d
from ..networks.sams import SamsGenerator
from .base_model import BaseModel


class SamsModel(BaseModel):
    """ Self Attentive Multi-Spade """

    @classmethod
    def modify_commandline_options(cls, parser: argparse.ArgumentParser, is_train):
        parser = argparse.ArgumentParser(parents=[parser], add_help=False)
        parser = super(SamsModel, cls).modify_commandline_options(parser, is_train)
        parser.set_defaults(person_inputs=("agnostic", "densepose", "flow"))
        parser.add_argument("--encoder_input", default="flow",
                            help="which of the --person_inputs to use as the encoder segmap input (only 1 allowed).")
        parser.set_defaults(n_frames_total=5)  # previous frames fed to the encoder = n_frames_total - 1
        parser.set_defaults(batch_size=4)
        for name in ("l1", "vgg", "multiscale", "temporal"):
            parser.add_argument(f"--wt_{name}", type=float, default=1.0, help=f"Weight of the {name} loss in the generator")
        parser.add_argument("--norm_D", type=str, default="spectralinstance")
        return SamsGenerator.modify_commandline_options(parser, is_train)

    @staticmethod
    def apply_default_encoder_input(opt):
        if hasattr(opt, "encoder_input") and opt.encoder_input is None:
            opt.encoder_input = opt.person_inputs[0]
        return opt

    def __init__(self, hparams):
        if isinstance(hparams, dict):
            hparams = argparse.Namespace(**hparams)
        super().__init__(hparams)
        self.n_frames_now = getattr(hparams, "n_frames_now", None) or self.n_frames_total
        self.inputs = list(hparams.person_inputs) + list(hparams.cloth_inputs)
        self.generator = SamsGenerator(hparams)
        self.resample = Resample2d()
        if self.is_train:
            raise NotImplementedError("SamsModel training (multiscale / temporal discriminators, GAN losses) is outside the "
                                      "hot path this package implements (SURVEY.md §2); build it with is_train=False")

    def set_precision(self, precision):
        self.generator.precision = precision

    def forward(self, *args, **kwargs):
        return self.generator(*args, **kwargs)

    def get_prev_frames_and_maps(self, batch, fIdx, all_G_frames):
        """sams_model.py:240-271: the ring window of previously generated frames and the encoder label maps of the frames
        before fIdx, zero-padded at the front."""
        enc = batch[self.hparams.encoder_input]
        n = self.n_frames_total
        if n == 1:
            return torch.zeros_like(all_G_frames), torch.zeros_like(enc)
        n_prev = n - 1
        idx = torch.tensor([(i + 1) % n for i in range(fIdx, fIdx + n_prev)], device=all_G_frames.device)
        prev = torch.index_select(all_G_frames, 1, idx)
        b, _, c, h, w = enc.shape
        start = n_prev - fIdx
        maps = torch.cat((enc.new_zeros(b, start, c, h, w), enc[:, start:-1]), dim=1)
        return prev, maps

    @torch.no_grad()
    def generate_n_frames(self, batch):
        """sams_model.py:204-238.  batch: {key: [b, n, c, h, w]} f32 CUDA tensors.  Returns (last frame, the label maps of
        the last frame, all generated frames [b, n, 3, h, w])."""
        labelmap = {k: batch[k] for k in self.inputs}
        frames = torch.zeros_like(batch["image"])
        flow_warp = bool(self.hparams.flow_warp)
        maps, fake = None, None
        for f in range(self.n_frames_total - self.n_frames_now, self.n_frames_total):
            maps = {k: v[:, f].contiguous() for k, v in labelmap.items()}
            prev, prev_maps = self.get_prev_frames_and_maps(batch, f, frames)
            out = self.generator.forward_nhwc(prev, prev_maps, maps)
            warped = None
            if flow_warp:
                last = frames[:, f - 1].contiguous() if f > 0 else torch.zeros_like(frames[:, f]).contiguous()
                warped = self.resample(last, batch["flow"][:, f].contiguous())
            fake = ops.sams_flow_blend(out, frames[:, f], warped)
        return fake, maps, frames

    def training_step(self, batch, batch_idx, optimizer_idx=0, val=False):
        raise NotImplementedError("SamsModel.training_step (GAN losses, discriminators) is out of scope (SURVEY.md §2)")

    def test_step(self, batch, batch_idx):
        pass  # sams_model.py:346-347
