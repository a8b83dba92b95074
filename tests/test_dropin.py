"""Drop-in surface of the class API and of the three legacy extension modules.

* FeatureL2Norm.forward / FeatureCorrelation.forward (warp.py:39-67) and UnetSkipConnectionBlock.forward
  (unet.py:188-198) used on their own, against the oracle's restatement of those methods.
* legacy_shims: `resample2d_cuda` / `channelnorm_cuda` / `correlation_cuda` module names with the reference's pybind
  signatures; here (build container) the reference's UNTOUCHED wrapper files import against them, on the GPU box the
  shim functions take the very argument lists the reference-built extensions take and give the same results.
"""
import inspect
import os
import sys

import pytest
import torch

from oracle import cases, flow_ops as fo, gmm, unet
from tests.util import assert_close, build_model

REF = "/root/reference"


# ------------------------------------------------------------------------------------------ CPU (build container)
def test_shim_modules_have_the_pybind_surface():
    from shineon_virtual_tryon_b200 import legacy_shims

    mods = legacy_shims.install()
    # argument counts of the pybind functions: resample2d_cuda.cc:6-31, channelnorm_cuda.cc:6-30, correlation_cuda.cc:10,89
    arity = {"resample2d_cuda": (5, 7), "channelnorm_cuda": (3, 5), "correlation_cuda": (11, 13)}
    for name, (nf, nb) in arity.items():
        m = sys.modules[name]
        assert m is mods[name]
        assert len(inspect.signature(m.forward).parameters) == nf, name
        assert len(inspect.signature(m.backward).parameters) == nb, name
    with pytest.raises(RuntimeError):  # CPU tensors are refused loudly (no fallback)
        sys.modules["channelnorm_cuda"].forward(torch.zeros(1, 3, 4, 4), torch.zeros(1, 1, 4, 4), 2)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference checkout only exists in the build container")
def test_reference_wrappers_import_against_the_shims():
    """The reference's own resample2d.py / channelnorm.py / correlation.py, unmodified, bind to the shim modules."""
    from oracle import ref_shim
    from shineon_virtual_tryon_b200 import legacy_shims

    ref_shim.install()
    legacy_shims.install(force=True)
    for mod in [m for m in list(sys.modules) if m.startswith("models.flownet2_pytorch.networks.") and m.endswith(("resample2d", "channelnorm", "correlation"))]:
        del sys.modules[mod]  # re-import so the module-level `import *_cuda` binds to the shims
    from models.flownet2_pytorch.networks.channelnorm_package import channelnorm as cn
    from models.flownet2_pytorch.networks.correlation_package import correlation as co
    from models.flownet2_pytorch.networks.resample2d_package import resample2d as rs

    assert rs.resample2d_cuda.forward is legacy_shims.resample2d_forward
    assert cn.channelnorm_cuda.backward is legacy_shims.channelnorm_backward
    assert co.correlation_cuda.forward is legacy_shims.correlation_forward
    assert rs.Resample2d(1, True) is not None and cn.ChannelNorm() is not None and co.Correlation(20, 1, 20, 1, 2, 1) is not None


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_reference_wrapper_call_sequences_through_the_shims(cuda):
    """The exact call sequences of the reference's autograd Functions (resample2d.py:8-40, channelnorm.py:8-31,
    correlation.py:9-45: caller-allocated zeroed outputs, empty tensors for correlation) through the shim modules,
    against the oracle, and — where oracle/_ref was built — against the reference's own extension called identically."""
    from oracle import build_ref
    from shineon_virtual_tryon_b200 import legacy_shims

    mods = legacy_shims.install()
    g = torch.Generator().manual_seed(31)
    img, flow = torch.rand(2, 3, 64, 48, generator=g), torch.randn(2, 2, 64, 48, generator=g) * 4
    go = torch.randn(2, 3, 64, 48, generator=g)

    def resample_calls(ext):
        out = img.cuda().new(2, 3, 64, 48).zero_()
        assert ext.forward(img.cuda(), flow.cuda(), out, 1, True) == 1
        g1, g2 = torch.zeros(2, 3, 64, 48, device="cuda"), torch.zeros(2, 2, 64, 48, device="cuda")
        assert ext.backward(img.cuda(), flow.cuda(), go.cuda(), g1, g2, 1, True) == 1
        return out, g1, g2

    ours = resample_calls(mods["resample2d_cuda"])
    assert_close(ours[0], fo.resample2d_fwd(img, flow), atol=1e-6, rtol=1e-5, what="resample2d fwd")
    w1, w2 = fo.resample2d_bwd(img, flow, go)
    assert_close(ours[1], w1, atol=1e-5, rtol=1e-4, what="resample2d d_in1")
    assert_close(ours[2], w2, atol=1e-5, rtol=1e-4, what="resample2d d_flow")
    ref = build_ref.load("resample2d_cuda")
    if ref is not None:
        for a, b in zip(ours, resample_calls(ref)):
            assert_close(a, b, atol=1e-5, rtol=1e-4, what="resample2d shim vs reference extension")

    x = torch.randn(2, 3, 40, 28, generator=g)
    out = torch.zeros(2, 1, 40, 28, device="cuda")
    assert mods["channelnorm_cuda"].forward(x.cuda(), out, 2) == 1
    assert_close(out, fo.channelnorm_fwd(x), atol=1e-6, rtol=1e-6, what="channelnorm fwd")
    gi = torch.zeros(2, 3, 40, 28, device="cuda")
    gout = torch.randn(2, 1, 40, 28, generator=g)
    assert mods["channelnorm_cuda"].backward(x.cuda(), out, gout.cuda(), gi, 2) == 1
    assert_close(gi, fo.channelnorm_bwd(x, out.cpu(), gout), atol=1e-6, rtol=1e-5, what="channelnorm bwd")

    a, b = torch.randn(2, 16, 12, 10, generator=g), torch.randn(2, 16, 12, 10, generator=g)
    ac, bc = a.cuda(), b.cuda()
    rb1, rb2, o = ac.new(), bc.new(), ac.new()
    assert mods["correlation_cuda"].forward(ac, bc, rb1, rb2, o, 4, 1, 4, 1, 2, 1) == 1
    assert tuple(o.shape) == (2,) + fo.correlation_out_shape(16, 12, 10, 4, 1, 4, 1, 2)
    assert_close(o, fo.correlation_fwd(a, b, 4, 1, 4, 1, 2), atol=1e-5, rtol=1e-4, what="correlation fwd")
    gout = torch.randn(o.shape, generator=g)
    g1, g2 = ac.new(), bc.new()
    assert mods["correlation_cuda"].backward(ac, bc, rb1, rb2, gout.cuda(), g1, g2, 4, 1, 4, 1, 2, 1) == 1
    w1, w2 = fo.correlation_bwd(a, b, gout, 4, 1, 4, 1, 2)
    assert_close(g1, w1, atol=1e-5, rtol=1e-4, what="correlation d_in1")
    assert_close(g2, w2, atol=1e-5, rtol=1e-4, what="correlation d_in2")


@pytest.mark.gpu
def test_feature_l2norm_and_correlation_forwards(cuda):
    from shineon_virtual_tryon_b200.networks.cpvton.warp import FeatureCorrelation, FeatureL2Norm

    g = torch.Generator().manual_seed(32)
    for (B, C, h, w) in [(2, 512, 16, 12), (3, 64, 5, 7)]:
        fa, fb = torch.randn(B, C, h, w, generator=g), torch.randn(B, C, h, w, generator=g)
        na = FeatureL2Norm()(fa.cuda())
        assert_close(na, gmm.feature_l2norm(fa), atol=1e-6, rtol=1e-5, what="FeatureL2Norm.forward")
        corr = FeatureCorrelation()(na, FeatureL2Norm()(fb.cuda()))
        want = gmm.feature_correlation(gmm.feature_l2norm(fa), gmm.feature_l2norm(fb))
        assert corr.shape == want.shape == (B, h * w, h, w)
        assert_close(corr, want, atol=1e-5, rtol=1e-4, what="FeatureCorrelation.forward")
        raw = FeatureCorrelation()(fa.cuda(), fb.cuda())  # un-normalised inputs: plain dot products
        assert_close(raw, gmm.feature_correlation(fa, fb), atol=1e-3, rtol=1e-4, what="FeatureCorrelation.forward (raw)")


@pytest.mark.gpu
@pytest.mark.parametrize("act,level", [("gelu", 5), ("gelu", 4), ("gelu", 2), (None, 3), ("gelu", 0)])
def test_unet_skip_block_forward_standalone(cuda, act, level):
    """A UnetSkipConnectionBlock called on its own: torch.cat([x, model(x)], 1) (unet.py:188-198), including the
    in-place LeakyReLU side effect of the default activation, vs the oracle's restatement of the same method."""
    over = dict(self_attn=True, num_attn=2, activation=act, ngf=64)
    model, sd = build_model("unet_mask", **over)
    blk, prefix = model.unet.model, "unet.model.model."
    for _ in range(level):
        idx = [i for i, m in enumerate(blk.model) if m is blk._parts["sub"]][0]
        prefix += f"{idx}.model."
        blk = blk._parts["sub"]
    cin = blk._parts["downconv"].in_channels
    H, W = 256 >> level, 192 >> level
    g = torch.Generator().manual_seed(33 + level)
    x = torch.randn(2, cin, H, W, generator=g)
    xc = x.clone().cuda()
    with torch.no_grad():
        got = blk(xc)
        want = unet._block(sd, prefix, x.clone(), level, 6, unet.attention_levels(6, 2, True), act)
    assert got.shape == want.shape
    assert_close(got, want, what=f"UnetSkipConnectionBlock.forward level {level} act {act}")
    if act is None and level > 0:  # the caller's tensor was rewritten in place like nn.LeakyReLU(0.2, True) does
        assert_close(xc, torch.nn.functional.leaky_relu(x, 0.2), atol=0, rtol=0, what="in-place side effect")
