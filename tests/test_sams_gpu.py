"""SAMS generator on cuda:0 against the CPU oracle (full resolution) and the golden vectors made by the
reference's own SamsGenerator (tests/golden/sams_*.npz).  Tolerance = north_star: abs <= 1e-3 OR rel <= 1e-2."""
import argparse

import pytest
import torch

from oracle import cases, sams as osams, weights
from tests.golden_util import load_golden
from tests.util import assert_close, nchw, nhwc

pytestmark = pytest.mark.gpu


def _hp(name):
    return argparse.Namespace(**cases.SAMS_CASES[name][0])


def _build(name):
    from shineon_virtual_tryon_b200.networks.sams import SamsGenerator

    seed, shapes, gold = load_golden(name)
    sd = weights.fix_spectral(weights.synth_state_dict(shapes, seed))
    g = SamsGenerator(_hp(name))
    g.load_state_dict(sd, strict=True)
    return g.cuda().eval(), sd, gold


@pytest.mark.parametrize("name", list(cases.SAMS_CASES))
def test_sams_generator_matches_oracle_and_golden(cuda, name):
    g, sd, gold = _build(name)
    prev, prev_maps, maps = cases.sams_inputs(name)
    c = lambda t: None if t is None else t.cuda()
    with torch.no_grad():
        got = g(c(prev), c(prev_maps), {k: v.cuda() for k, v in maps.items()})
        want = osams.generator_forward(sd, _hp(name), prev, prev_maps, maps)
    torch.cuda.synchronize()
    err = assert_close(got, want, what=f"{name} vs oracle")
    H = cases.SAMS_CASES[name][2]
    assert_close(cases.subsample(got.cpu(), 4 if H >= 256 else 1), gold["out"], what=f"{name} vs reference golden")
    print(f"{name}: max abs err {err:.2e} (output rms {want.pow(2).mean().sqrt().item():.3f})")


def test_sams_training_mode_raises(cuda):
    g, _, _ = _build("sams_small")
    prev, prev_maps, maps = cases.sams_inputs("sams_small")
    g.train()
    with pytest.raises(NotImplementedError):
        g(prev.cuda(), prev_maps.cuda(), {k: v.cuda() for k, v in maps.items()})


@pytest.mark.parametrize("norm_mode", ["instance", "batch", "none"])
@pytest.mark.parametrize("C,act", [(64, "gelu"), (20, "leaky"), (7, None)])
def test_spade_modulate_kernel(cuda, norm_mode, C, act):
    """One modulation pass against the torch fp32 expression of spade.py:68-84 (all vector widths, every norm flavour)."""
    import torch.nn.functional as F
    from shineon_virtual_tryon_b200 import ops

    g = torch.Generator().manual_seed(C)
    N, H, W = 3, 9, 7
    x = torch.randn(N, C, H, W, generator=g) * 2 + 0.5
    gamma, beta = torch.randn(N, C, H, W, generator=g), torch.randn(N, C, H, W, generator=g)
    rm, rv = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    if norm_mode == "instance":
        nrm = F.instance_norm(x, eps=1e-5)
    elif norm_mode == "batch":
        nrm = F.batch_norm(x, rm, rv, training=False, eps=1e-5)
    else:
        nrm = x
    want = nrm * (1 + gamma) + beta
    want = {"gelu": F.gelu, "leaky": lambda t: F.leaky_relu(t, 0.2), None: lambda t: t}[act](want)
    xd = nhwc(x).cuda()
    gb = torch.cat([nhwc(1 + gamma), nhwc(beta)], -1).contiguous().cuda()
    kw = {}
    if norm_mode == "instance":
        kw["stats_ws"] = ops.chan_stats(xd)
    elif norm_mode == "batch":
        rstd = torch.rsqrt(rv + 1e-5)
        kw.update(nscale=rstd.cuda(), nshift=(-rm * rstd).cuda())
    yf, yp = ops.spade_modulate(xd, gb, act=act, act_param=0.2, want_f32=True, want_planes=True, **kw)
    torch.cuda.synchronize()
    assert_close(nchw(yf), want, atol=2e-5, rtol=1e-5, what="f32 output")
    assert_close(yp.float(), want, atol=2e-5, rtol=1e-5, what="hi/lo planes")
    assert yp.hi[..., C:].abs().max().item() == 0 if yp.cpad > C else True


@pytest.mark.parametrize("factor", [0.5, 2])
def test_nearest_resize_matches_interpolate(cuda, factor):
    import torch.nn.functional as F
    from shineon_virtual_tryon_b200 import ops

    x = torch.randn(2, 12, 10, 6)
    got = ops.nearest_resize_nhwc(nhwc(x).cuda(), factor)
    assert torch.equal(nchw(got).cpu(), F.interpolate(x, scale_factor=factor, mode="nearest"))
    seg = torch.randn(2, 5, 16, 12)
    for size in ((8, 6), (4, 3), (16, 12), (32, 24)):
        p = ops.nearest_resize_planes(seg.cuda(), size)
        want = F.interpolate(seg, size=size, mode="nearest")
        assert_close(p.float(), want, atol=1e-6, rtol=1e-6, what=f"planes {size}")  # hi + lo carries 22 mantissa bits


def test_sams_model_generate_n_frames(cuda):
    """SamsModel.generate_n_frames (3-frame window, flow blend through Resample2d) against the oracle and the golden made by
    the reference's own SamsModel; the frame buffer feeds back into the encoder, so errors would compound."""
    from shineon_virtual_tryon_b200.models import find_model_using_name
    from oracle import flow_ops as fo

    name = "sams_small"
    seed, shapes, gold = load_golden(name + "_model")
    sd = weights.fix_spectral(weights.synth_state_dict(shapes, seed))
    hp = argparse.Namespace(**cases.SAMS_CASES[name][0], is_train=False)
    m = find_model_using_name("sams")(hp)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    batch = cases.sams_model_batch(name)
    last, maps, frames = m.generate_n_frames({k: v.cuda() for k, v in batch.items()})
    torch.cuda.synchronize()
    gsd = {k[len("generator."):]: v for k, v in sd.items() if k.startswith("generator.")}
    with torch.no_grad():
        want_last, want_frames = osams.generate_n_frames(gsd, hp, batch, fo.resample2d_fwd)
    assert sorted(maps) == sorted(hp.person_inputs + hp.cloth_inputs)
    err = assert_close(frames, want_frames, what="frames vs oracle")
    assert_close(last, want_last, what="last frame vs oracle")
    assert_close(frames.cpu(), gold["frames"], what="frames vs reference golden")
    assert_close(last.cpu(), gold["last"], what="last frame vs reference golden")
    print(f"generate_n_frames: max abs err {err:.2e}")
