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
        self.is_train = True
        return parser
