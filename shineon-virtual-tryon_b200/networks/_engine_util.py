"""Shared helpers of the host-side engines: packed-weight cache invalidation, BN folding, side-stream branches."""
import os

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


# ------------------------------------------------------------------ side-stream branches of the backward pass
# A layer's weight / bias gradient (conv_wgrad + its split reduction + channel_sum) does not feed the rest of the backward:
# it runs on an auxiliary stream next to the input-gradient chain (a parallel branch of the captured training graph).  At
# the recipe's batch of 4 per GPU both chains are made of kernels that leave most SMs idle.  Callers keep every tensor the
# branch reads alive until they have called the returned join (which makes the current stream wait for the branch), so
# the caching allocator cannot hand those blocks to later kernels early.
SIDE_WGRAD = os.environ.get("SHINEON_WGRAD_PARALLEL", "1") != "0"
_AUX_STREAMS = {}


def _no_join():
    return None


def on_aux_stream(device):
    """True while the current stream of `device` is side_run's auxiliary stream."""
    aux = _AUX_STREAMS.get(torch.device(device))
    return aux is not None and torch.cuda.current_stream(device) == aux


def side_run(fn, enabled=True):
    """fn() on the auxiliary stream of the current device, forked from the current stream; returns join()."""
    from .. import ops

    if not (SIDE_WGRAD and enabled) or ops.PROFILE is not None:
        fn()
        return _no_join
    cur = torch.cuda.current_stream()
    aux = _AUX_STREAMS.get(cur.device)
    if aux is None:
        aux = _AUX_STREAMS[cur.device] = torch.cuda.Stream(cur.device)
    if aux == cur:  # already on the auxiliary stream (nested branch): stay serial
        fn()
        return _no_join
    aux.wait_stream(cur)
    with torch.cuda.stream(aux):
        fn()
    return lambda: cur.wait_stream(aux)
