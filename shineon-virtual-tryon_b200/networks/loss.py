"""Losses of the reference's models/networks/loss.py that sit on the hot path: VGGLoss (loss.py:106-122), used by
UnetMaskModel.training_step.  GANLoss belongs to the SAMS-GAN model (out of scope)."""
from .vgg import VGGLoss, Vgg19  # noqa: F401
