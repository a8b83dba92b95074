"""Resample2d / ChannelNorm / Correlation — same Module + autograd.Function surface as the reference's three
CUDA extensions (resample2d.py:5-49, channelnorm.py:5-38, correlation.py:6-60), backed by libshineon_b200.so.

Differences kept on purpose: outputs are allocated by the wrapper (the reference passes empty tensors that
its C++ side resizes, correlation.py:20-25), no padded NHWC scratch tensors (rbot1/rbot2) exist, and a failed
launch raises instead of being swallowed.
"""
import torch
from torch.autograd import Function
from torch.nn.modules.module import Module

from ... import ops


class Resample2dFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2, kernel_size=1, bilinear=True):
        assert input1.is_contiguous()
        assert input2.is_contiguous()
        ctx.save_for_backward(input1, input2)
        ctx.kernel_size = kernel_size
        ctx.bilinear = bilinear
        return ops.resample2d_fwd(input1, input2, kernel_size, bilinear)

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        g1, g2 = ops.resample2d_bwd(input1, input2, grad_output.contiguous(), ctx.kernel_size, ctx.bilinear)
        return g1, g2, None, None


class Resample2d(Module):
    def __init__(self, kernel_size=1, bilinear=True):
        super().__init__()
        self.kernel_size = kernel_size
        self.bilinear = bilinear

    def forward(self, input1, input2):
        return Resample2dFunction.apply(input1.contiguous(), input2, self.kernel_size, self.bilinear)


class ChannelNormFunction(Function):
    @staticmethod
    def forward(ctx, input1, norm_deg=2):
        assert input1.is_contiguous()
        output = ops.channelnorm_fwd(input1, norm_deg)
        ctx.save_for_backward(input1, output)
        ctx.norm_deg = norm_deg
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input1, output = ctx.saved_tensors
        return ops.channelnorm_bwd(input1, output, grad_output.contiguous(), ctx.norm_deg), None


class ChannelNorm(Module):
    def __init__(self, norm_deg=2):
        super().__init__()
        self.norm_deg = norm_deg

    def forward(self, input1):
        return ChannelNormFunction.apply(input1, self.norm_deg)


class CorrelationFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2, pad_size=3, kernel_size=3, max_displacement=20, stride1=1, stride2=2,
                corr_multiply=1):
        input1, input2 = input1.contiguous(), input2.contiguous()
        ctx.save_for_backward(input1, input2)
        ctx.cfg = (pad_size, kernel_size, max_displacement, stride1, stride2)
        return ops.correlation_fwd(input1, input2, *ctx.cfg)

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        g1, g2 = ops.correlation_bwd(input1, input2, grad_output.contiguous(), *ctx.cfg)
        return g1, g2, None, None, None, None, None, None


class Correlation(Module):
    def __init__(self, pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1):
        super().__init__()
        self.pad_size = pad_size
        self.kernel_size = kernel_size
        self.max_displacement = max_displacement
        self.stride1 = stride1
        self.stride2 = stride2
        self.corr_multiply = corr_multiply

    def forward(self, input1, input2):
        return CorrelationFunction.apply(input1, input2, self.pad_size, self.kernel_size, self.max_displacement,
                                         self.stride1, self.stride2, self.corr_multiply)
