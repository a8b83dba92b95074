"""Import shim for the UNMODIFIED reference checkout (build container only; /root/reference does not
exist on the GPU box).  Used by oracle/make_golden.py and, when present, by bench.py --impl reference.

The reference does not import on Python 3.12 / without its pinned deps for four reasons, each handled
here without touching the reference tree:
  1. util/__init__.py:2 `from collections import Iterable`        -> alias collections.Iterable
  2. pytorch_lightning 0.9 is absent (base_model.py:9 ...)         -> stub module, LightningModule = nn.Module
  3. matplotlib / colorlog / pytz absent                           -> empty stub modules
  4. the three CUDA extensions are imported at module top          -> stub modules (their ops are replaced by
     (resample2d.py:3, correlation.py:4, channelnorm.py:3)            the CPU oracle when a forward needs them)
"""
import collections
import collections.abc
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SHINEON_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Make `import models...` resolve to the reference.  Idempotent."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    if getattr(install, "_done", False):
        return
    import torch
    from torch import nn

    sys.dont_write_bytecode = True
    collections.Iterable = collections.abc.Iterable

    class _Result(dict):
        def __init__(self, *a, **k):
            super().__init__()
            self.minimize = a[0] if a else k.get("minimize")

        def log(self, name, value, **kw):
            self[name] = value

    class _LightningModule(nn.Module):
        global_step = 0
        logger = None

    pl = _stub("pytorch_lightning", LightningModule=_LightningModule, TrainResult=_Result, EvalResult=_Result,
               Trainer=object, Callback=object)
    cb = _stub("pytorch_lightning.callbacks", ModelCheckpoint=object, Callback=object)
    pl.callbacks = cb
    for name in ("matplotlib", "matplotlib.pyplot", "colorlog", "pytz", "matplotlib.colors", "matplotlib.cm"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:  # noqa: BLE001
                _stub(name)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules.get("matplotlib.pyplot")
    for name in ("resample2d_cuda", "correlation_cuda", "channelnorm_cuda"):
        if name not in sys.modules:
            _stub(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    install._done = True


def patch_native_ops_with_oracle():
    """Route the reference's three CUDA-only autograd Functions to the CPU oracle so FlowNet2 /
    UnetMaskModel(flows=...) forwards can run on CPU."""
    install()
    import torch
    from oracle import flow_ops as fo
    from models.flownet2_pytorch.networks.resample2d_package import resample2d as r2
    from models.flownet2_pytorch.networks.channelnorm_package import channelnorm as cn
    from models.flownet2_pytorch.networks.correlation_package import correlation as co

    r2.Resample2d.forward = lambda self, a, b: fo.resample2d_fwd(a.contiguous(), b, self.kernel_size, self.bilinear)
    cn.ChannelNorm.forward = lambda self, a: fo.channelnorm_fwd(a)
    co.Correlation.forward = lambda self, a, b: fo.correlation_fwd(a, b, self.pad_size, self.kernel_size,
                                                                  self.max_displacement, self.stride1, self.stride2)


def hparams(**over):
    """A minimal option Namespace with the fields the reference model ctors read."""
    import argparse

    d = dict(n_frames_total=1, n_frames_now=1, person_inputs=["agnostic", "densepose"], cloth_inputs=["cloth"],
             ngf=64, self_attn=True, num_attn=2, flow_warp=False, activation="gelu", is_train=False, grid_size=5,
             fine_height=256, fine_width=192, pen_flow_mask=1.0, display_count=1000000, lr=1e-4)
    d.update(over)
    return argparse.Namespace(**d)


def build_warp_model(**over):
    install()
    from models.warp_model import WarpModel

    hp = hparams(person_inputs=["agnostic", "cocopose"], **over)
    return WarpModel(hp).eval()


def build_unet_mask_model(**over):
    install()
    import torch
    from torch import nn
    import models.unet_mask_model as umm

    # VGGLoss() downloads torchvision's ImageNet weights and calls .cuda() (loss.py:106-110, vgg.py:9): build the same
    # torchvision architecture without the download and keep it on the CPU — the reference classes stay untouched
    import torchvision.models as tvm
    import models.networks.vgg as ref_vgg

    class _TV:
        @staticmethod
        def vgg19(pretrained=False, **kw):
            return tvm.vgg19(weights=None)

    ref_vgg.models = _TV
    orig_cuda = nn.Module.cuda
    nn.Module.cuda = lambda self, *a, **k: self
    try:
        m = umm.UnetMaskModel(hparams(**over))
    finally:
        nn.Module.cuda = orig_cuda
    return m.eval()
