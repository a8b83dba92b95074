"""Sine / Swish activation modules (reference: models/networks/activation.py:4-18).

Inside the U-Net engine these are fused into the producing kernels (conv epilogue / InstanceNorm /
attention epilogue, csrc/common.cuh:apply_act); the modules exist so the module tree and its state_dict
match the reference.  Called standalone they are plain pointwise torch expressions.
"""
import torch
from torch import nn


class Sine(nn.Module):
    def forward(self, input):
        return torch.sin(30 * input)


class Swish(nn.Module):
    def forward(self, input_tensor):
        return input_tensor * torch.sigmoid(input_tensor)


def act_name(module):
    """Kernel activation id (and parameter) for an activation module of the reference's U-Net."""
    if module is None:
        return None, 0.0
    if isinstance(module, nn.LeakyReLU):
        return "leaky", float(module.negative_slope)
    if isinstance(module, nn.ReLU):
        return "relu", 0.0
    if isinstance(module, nn.GELU):
        return "gelu", 0.0
    if isinstance(module, Swish):
        return "swish", 0.0
    if isinstance(module, Sine):
        return "sine", 0.0
    raise NotImplementedError(f"no fused kernel for activation {type(module).__name__}")
