"""TrainOptions (options/train_options.py:6-52)."""
from .base_options import BaseOptions


class TrainOptions(BaseOptions):
    def initialize(self, parser):
        parser = BaseOptions.initialize(self, parser)
        parser.add_argument("--no_shuffle", action="store_true", help="don't shuffle input data")
        parser.add_argument("--save_count", type=int, default=10000, help="how often in steps to always save a checkpoint")
        parser.add_argument("--val_check_interval", "--val_frequency", dest="val_check_interval", type=str, default="0.125")
        parser.add_argument("--lr", type=float, default=1e-4, help="initial learning rate for adam")
        parser.add_argument("--keep_epochs", type=int, default=5, help="number of epochs with initial learning rate")
        parser.add_argument("--decay_epochs", type=int, default=5, help="number of epochs to linearly decay the learning rate")
        parser.add_argument("--accumulated_batches", type=int, default=1,
                            help="number of batch gradients to accumulate before calling optimizer.step()")
        parser.add_argument("--b200_train_precision", choices=("bf16x3", "bf16"), default="bf16",
                            help="numeric mode of the native training step: bf16 = single bf16 tensor-core products with "
                                 "fp32 accumulation / statistics / master weights (the reference's default is AMP fp16); "
                                 "bf16x3 = fp32-grade split products (the parity-tested mode)")
        parser.add_argument("--vgg_weights", default=None,
                            help="torchvision vgg19 state_dict file for the perceptual loss (the reference downloads the "
                                 "ImageNet weights, models/networks/vgg.py:9; without them the VGG slices are random and "
                                 "training warns)")
        parser.add_argument("--max_steps", type=int, default=None, help="stop after this many optimiser steps")
        self.is_train = True
        return parser
