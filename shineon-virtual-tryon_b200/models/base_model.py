"""BaseModel — the model-side API surface of the reference's LightningModule base (models/base_model.py:24-237)
without the Lightning dependency: option flags, hparams handling, optimiser/scheduler construction and the
step hooks.  Dataset / dataloader construction is the reference's CPU data path and is
injected by the caller (`train_dataset` / `val_dataset` attributes) rather than re-implemented here.
"""
import abc
import argparse
import os.path as osp

import torch
from torch import nn
from torch.optim import Adam

# channel table of datasets/tryon_dataset.py:47-61 (the contract for input tensor shapes)
CHANNELS = dict(RGB=3, MASK=1, COCOPOSE=18, IM_HEAD=3, SILHOUETTE=1, AGNOSTIC=4, CLOTH=3, CLOTH_MASK=1,
                DENSEPOSE=3, FLOW=2, IMAGE=3, PREV_IMAGE=3)


def parse_num_channels(list_of_inputs):
    """datasets/tryon_dataset.py:540-547."""
    if isinstance(list_of_inputs, str):
        list_of_inputs = [list_of_inputs]
    return sum(CHANNELS[inp.upper()] for inp in list_of_inputs)


class BaseModel(nn.Module, abc.ABC):
    @classmethod
    def modify_commandline_options(cls, parser: argparse.ArgumentParser, is_train):
        parser.add_argument("--person_inputs", nargs="+",
                            help="List of what type of items are passed as person input.")
        parser.add_argument("--cloth_inputs", nargs="+", default=("cloth",),
                            help="List of items to pass as the cloth inputs.")
        parser.add_argument("--ngf", type=int, default=64)
        parser.add_argument("--self_attn", action="store_true", help="Add self-attention")
        parser.add_argument("--no_self_attn", action="store_false", dest="self_attn", help="No self-attention")
        parser.add_argument("--num_attn", type=int, default=2,
                            help="Num of self-attention layers: start layers from bottom of UNet all the way up the U")
        parser.add_argument("--flow_warp", action="store_true", help="Warp the previous frame with flow")
        return parser

    def __init__(self, hparams, *args, **kwargs):
        if isinstance(hparams, dict):
            hparams = argparse.Namespace(**hparams)
        super().__init__(*args, **kwargs)
        self.hparams = hparams
        self.n_frames_total = hparams.n_frames_total
        self.person_channels = parse_num_channels(hparams.person_inputs)
        self.cloth_channels = parse_num_channels(hparams.cloth_inputs)
        self.is_train = self.hparams.is_train
        self.global_step = 0
        if self.is_train:
            self.val_visualization_batch = None

    def override_hparams(self, hparams: argparse.Namespace):
        """Re-apply non-architectural hparams after a checkpoint load (base_model.py:76-89)."""
        self.hparams = hparams
        if not self.is_train:
            ckpt_name = osp.basename(hparams.checkpoint)
            self.test_results_dir = osp.join(hparams.result_dir, hparams.name, ckpt_name, hparams.datamode)

    def validation_step(self, batch, idx):
        self.val_visualization_batch = batch
        return self.training_step(batch, idx, val=True)

    def visualize(self, input_batch, tag="train"):
        pass

    def configure_optimizers(self):
        optimizer = Adam(self.parameters(), self.hparams.lr)
        scheduler = self._make_step_scheduler(optimizer)
        return [optimizer], [scheduler]

    def _make_step_scheduler(self, optimizer):
        def step_func(epoch):
            decrease = max(0, epoch - self.hparams.keep_epochs) / float(self.hparams.decay_epochs + 1)
            return 1.0 - decrease

        return torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda=step_func)


def get_and_cat_inputs(batch, names):
    """util/__init__.py:64-66."""
    return torch.cat([batch[inp] for inp in names], dim=1)


def maybe_combine_frames_and_channels(opt, inputs):
    """datasets/n_frames_interface.py:105-138: fold [B,N,C,H,W] -> [B,N*C,H,W] for every 5-D tensor."""
    if not hasattr(opt, "n_frames_total"):
        return inputs
    out = {}
    for k, v in inputs.items():
        if isinstance(v, torch.Tensor) and v.dim() == 5:
            b, n, c, h, w = v.shape
            out[k] = v.reshape(b, n * c, h, w)
        else:
            out[k] = v
    return out
