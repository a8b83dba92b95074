"""BaseOptions — same flags, defaults, synonyms and post-parse normalisation as the reference
(options/base_options.py:13-265).  Differences: dataset-specific flags come from the synthetic / injected dataset
contract below (the reference's dataset classes are its CPU data path, out of scope here), and the interactive
"experiment name" prompt (base_options.py:199-211) only fires on a TTY."""
import argparse
import sys

from .. import models


def dataset_options(parser, is_train):
    """Flags the hot path reads from the dataset side (datasets/tryon_dataset.py:64-96, n_frames_interface.py:34-60)."""
    parser.add_argument("--fine_width", type=int, default=192)
    parser.add_argument("--fine_height", type=int, default=256)
    parser.add_argument("--n_frames_total", type=int, default=1, help="frames the model sees per sample")
    parser.add_argument("--n_frames_now", type=int, default=None, help="defaults to n_frames_total")
    parser.add_argument("--visualize_flow", action="store_true")
    parser.add_argument("--synthetic_samples", type=int, default=16, help="size of the built-in synthetic dataset")
    return parser


class BaseOptions:
    def __init__(self):
        self.initialized = False
        self.is_train = None

    def initialize(self, parser):
        parser.add_argument("--name", default="unnamed_experiment")
        parser.add_argument("--distributed_backend", default="ddp", help="how to do distributed multigpu training")
        parser.add_argument("--gpu_ids", default="0", help="comma separated of which GPUs to train on")
        parser.add_argument("-j", "--num_workers", "--workers", dest="workers", type=int, default=4)
        parser.add_argument("-b", "--batch_size", type=int, default=8)
        parser.add_argument("--activation", choices=("relu", "gelu", "swish", "sine"))
        parser.add_argument("-fp", "--precision", type=int, dest="precision", choices=(16, 32), default=16,
                            help="reference flag (AMP fp16 / fp32); the B200 numeric mode is --b200_precision")
        parser.add_argument("--b200_precision", choices=("fp16x3", "bf16x3", "fp16", "bf16"), default="fp16x3",
                            help="tensor-core numeric mode of this build (DESIGN.md section 4)")
        parser.add_argument("--dataset", choices=("viton", "viton_vvt_mpv", "vvt", "mpv", "synthetic"), default="synthetic")
        parser.add_argument("--datamode", default="train")
        parser.add_argument("--model", help="which model to use: 'warp' (aka 'gmm'), 'unet_mask' (aka 'tom').")
        parser.add_argument("--datacap", "--datacap_train", "--limit_train_batches", dest="limit_train_batches", default="1.0")
        parser.add_argument("--datacap_val", "--limit_val_batches", dest="limit_val_batches", default="1.0")
        parser.add_argument("--experiments_dir", default="experiments", help="where to store logs and checkpoints")
        parser.add_argument("--checkpoint", type=str, default="", help="model checkpoint for initialization")
        parser.add_argument("--display_count", type=int, default=200, help="how often to update tensorboard, in steps")
        parser.add_argument("--loglevel", choices=("debug", "info", "warning", "error", "critical"), default="info")
        parser.add_argument("--fast_dev_run", action="store_true", help="quickly test out the pipeline")
        self.initialized = True
        return parser

    def gather_options(self, argv=None):
        parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
        parser = self.initialize(parser)
        opt, _ = parser.parse_known_args(argv)
        if opt.model is None:
            parser.error("--model is required (warp | unet_mask, or the synonyms gmm | tom | unet)")
        BaseOptions.apply_model_synonyms(opt)
        parser = models.get_option_setter(opt.model)(parser, self.is_train)  # the model adds / overrides flags
        parser = dataset_options(parser, self.is_train)
        self.parser = parser
        return parser.parse_args(argv)

    def parse(self, argv=None):
        opt = self.gather_options(argv)
        opt.is_train = self.is_train
        BaseOptions.apply_ask_unnamed_experiment(opt, argv)
        BaseOptions.apply_model_synonyms(opt)
        BaseOptions.apply_gpu_ids(opt)
        BaseOptions.apply_sort_inputs(opt)
        if opt.n_frames_now is None:  # NFramesInterface.apply_n_frames_now_default_total
            opt.n_frames_now = opt.n_frames_total
        self.opt = opt
        return opt

    @staticmethod
    def apply_ask_unnamed_experiment(opt, argv=None):
        args = sys.argv if argv is None else argv
        if "--name" not in args and sys.stdin.isatty():
            new_name = input(f"Experiment name (default: {opt.name}): ")
            if new_name:
                opt.name = new_name

    @staticmethod
    def apply_gpu_ids(opt):
        opt.gpu_ids = [int(s) for s in str(opt.gpu_ids).split(",") if int(s) >= 0]

    @staticmethod
    def apply_model_synonyms(opt):
        opt.model = opt.model.lower()
        if opt.model == "gmm":
            opt.model = "warp"
        elif opt.model in ("tom", "unet"):
            opt.model = "unet_mask"

    @staticmethod
    def apply_sort_inputs(opt):
        opt.person_inputs = sorted(opt.person_inputs)
        opt.cloth_inputs = sorted(opt.cloth_inputs)
