"""SAMS generator: the CPU oracle against the golden vectors made by the reference's own SamsGenerator
(oracle/make_golden_sams.py), and the module tree of our SamsGenerator against the reference's state_dict keys.  CPU only."""
import argparse

import pytest
import torch

from oracle import cases, sams, weights
from tests.golden_util import load_golden
from tests.util import assert_close


def _hp(name):
    return argparse.Namespace(**cases.SAMS_CASES[name][0])


@pytest.mark.parametrize("name", list(cases.SAMS_CASES))
def test_sams_oracle_matches_reference_golden(name):
    seed, shapes, gold = load_golden(name)
    sd = weights.fix_spectral(weights.synth_state_dict(shapes, seed))
    prev, prev_maps, maps = cases.sams_inputs(name)
    with torch.no_grad():
        out = sams.generator_forward(sd, _hp(name), prev, prev_maps, maps)
    H = cases.SAMS_CASES[name][2]
    # same ATen kernels and graph as the reference; the tolerance covers thread-count effects and the power iteration
    assert_close(cases.subsample(out, 4 if H >= 256 else 1), gold["out"], atol=2e-5, rtol=1e-4, what="generator output")
    assert gold["out"].abs().mean() > 1e-2, "degenerate golden (all-zero output would hide everything)"


@pytest.mark.parametrize("name", list(cases.SAMS_CASES))
def test_sams_state_dict_matches_reference_key_for_key(name):
    from shineon_virtual_tryon_b200.networks.sams import SamsGenerator

    _, shapes, _ = load_golden(name)
    mine = {k: tuple(v.shape) for k, v in SamsGenerator(_hp(name)).state_dict().items()}
    assert mine == shapes


def test_spectral_eval_weight_is_normalised():
    """fix_spectral's (u, v) give a sigma = u . W v close below the top singular value (a few power iterations on a random
    matrix do not converge further), so weight_orig / sigma is O(1)-normalised like a trained checkpoint's."""
    _, shapes, _ = load_golden("sams_small")
    sd = weights.fix_spectral(weights.synth_state_dict(shapes, 420))
    k = "encode_layers.1.conv_1"
    w = sd[k + ".weight_orig"].reshape(sd[k + ".weight_orig"].shape[0], -1)
    sigma = torch.dot(sd[k + ".weight_u"], torch.mv(w, sd[k + ".weight_v"]))
    top = torch.linalg.matrix_norm(w, ord=2)
    assert 0.6 < float(sigma / top) <= 1.0 + 1e-5


def test_sams_model_oracle_matches_reference_golden():
    """generate_n_frames restated (oracle/sams.py) against SamsModel.generate_n_frames of the reference."""
    from oracle import flow_ops as fo

    name = "sams_small"
    seed, shapes, gold = load_golden(name + "_model")
    sd = weights.fix_spectral(weights.synth_state_dict(shapes, seed))
    gsd = {k[len("generator."):]: v for k, v in sd.items() if k.startswith("generator.")}
    with torch.no_grad():
        last, frames = sams.generate_n_frames(gsd, _hp(name), cases.sams_model_batch(name), fo.resample2d_fwd)
    assert_close(frames, gold["frames"], atol=2e-5, rtol=1e-4, what="frames")
    assert_close(last, gold["last"], atol=2e-5, rtol=1e-4, what="last")


def test_sams_model_state_dict_matches_reference():
    from shineon_virtual_tryon_b200.models import find_model_using_name

    _, shapes, _ = load_golden("sams_small_model")
    hp = argparse.Namespace(**cases.SAMS_CASES["sams_small"][0], is_train=False)
    mine = {k: tuple(v.shape) for k, v in find_model_using_name("sams")(hp).state_dict().items()}
    assert mine == shapes
