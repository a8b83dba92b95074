"""Shared helpers of the host-side engines: packed-weight cache invalidation and BN folding."""
import torch
from torch import nn


# Bumped by optimisers that update parameters through raw pointers (the fused Adam over the flat parameter
# buffer does not touch the tensors' autograd version counters): every packed-weight cache is keyed on it.
WEIGHTS_EPOCH = [0]


def bump_weights_epoch():
    WEIGHTS_EPOCH[0] += 1


def params_signature(module):
    """Cheap change detector for a module's parameters/buffers: (data_ptr, _version) of every tensor."""
    # frozen modules (the VGG19 slices of the perceptual loss) are never touched by the optimiser
    sig = [WEIGHTS_EPOCH[0] if any(p.requires_grad for p in module.parameters()) else -1]
    for t in list(module.parameters()) + list(module.buffers()):
        sig.append((t.data_ptr(), t._version, t.device.index))
    return tuple(sig)


def fold_bn(bn):
    """Eval-mode BatchNorm2d as per-channel (scale, shift) for the conv epilogue."""
    if bn.training:
        raise NotImplementedError(
            "BatchNorm2d in training mode (batch statistics) has no native kernel yet; call .eval() "
            "(the try-on inference path, test.py, always does)")
    w = bn.weight.detach().float() if bn.weight is not None else torch.ones_like(bn.running_var)
    b = bn.bias.detach().float() if bn.bias is not None else torch.zeros_like(bn.running_var)
    scale = w / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = b - bn.running_mean.detach().float() * scale
    return scale.contiguous(), shift.contiguous()


def require_cuda(module, what):
    p = next(module.parameters(), None)
    if p is None or not p.is_cuda:
        raise RuntimeError(f"{what}: module is not on a CUDA device; the shineon B200 path has no CPU fallback")
    return p.device
