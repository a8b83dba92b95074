"""Deterministic synthetic weights for the parity tests, golden fixtures and the benchmark.

No checkpoint of the reference is available offline, so tests use seeded random weights with the
reference's own init scales (models/networks/__init__.py:49-60: conv/linear N(0, 0.02), BN weight N(1, 0.02))
plus non-trivial values for everything that would otherwise hide a kernel: attention gamma (zero-initialised,
sagan.py:25), BatchNorm running statistics, biases.  A tensor's values depend only on (seed, key, shape),
so any process can regenerate exactly the same state_dict from the key/shape list.
"""
import zlib

import torch


def _gen(seed, key):
    return torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))


def synth_tensor(seed, key, shape, conv_std=None):
    g = _gen(seed, key)
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.long)
    if leaf == "gamma":
        return torch.rand(shape, generator=g) + 0.5
    if leaf == "running_mean":
        return torch.randn(shape, generator=g) * 0.1
    if leaf == "running_var":
        return torch.rand(shape, generator=g) + 0.5
    if leaf == "weight" and len(shape) == 1:  # BatchNorm affine
        return torch.rand(shape, generator=g) + 0.5
    if leaf == "bias":
        return torch.randn(shape, generator=g) * 0.05
    if leaf == "weight" and len(shape) == 4 and "criterionVGG" in key:
        # stands in for the ImageNet weights the reference downloads (vgg.py:9): fan-in scaled so the 13-layer ReLU
        # stack keeps O(1) activations and the perceptual term carries signal
        fan_in = shape[1] * shape[2] * shape[3]
        return torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
    if leaf in ("weight", "weight_orig") and len(shape) >= 2:
        # the reference's own initialiser scale: init.normal_(w, 0.0, 0.02) for Conv / Linear
        # (models/networks/__init__.py:49-55)
        std = 0.02 if conv_std is None else conv_std
        return torch.randn(shape, generator=g) * std
    return torch.randn(shape, generator=g) * 0.1


def synth_state_dict(shapes, seed=420, conv_std=None):
    """shapes: {key: shape}.  Returns {key: tensor}."""
    return {k: synth_tensor(seed, k, tuple(s), conv_std) for k, s in shapes.items()}


def fix_spectral(sd, iters=5):
    """Replace every synthesized (weight_u, weight_v) pair of a spectral_norm'd conv by the result of `iters` power
    iterations on its weight_orig (torch.nn.utils.spectral_norm.compute_weight's update, run on the CPU), so the
    eval-mode weight W / (u . W v) is a properly normalised one instead of W divided by a random number."""
    import torch.nn.functional as F

    for k in list(sd):
        if not k.endswith(".weight_orig"):
            continue
        base = k[: -len("weight_orig")]
        w = sd[k].reshape(sd[k].shape[0], -1).double()
        u = F.normalize(sd[base + "weight_u"].double().abs() + 0.1, dim=0)
        for _ in range(iters):
            v = F.normalize(torch.mv(w.t(), u), dim=0, eps=1e-12)
            u = F.normalize(torch.mv(w, v), dim=0, eps=1e-12)
        sd[base + "weight_u"], sd[base + "weight_v"] = u.float(), v.float()
    return sd


def shapes_of(module):
    return {k: tuple(v.shape) for k, v in module.state_dict().items()}
