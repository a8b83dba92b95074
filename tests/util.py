"""Shared helpers for the parity tests."""
import torch

# north_star tolerance for floating point: 1e-3 abs OR 1e-2 rel, element-wise
ATOL, RTOL = 1e-3, 1e-2


def assert_close(actual, expected, atol=ATOL, rtol=RTOL, what=""):
    actual = actual.detach().float().cpu()
    expected = expected.detach().float().cpu()
    assert actual.shape == expected.shape, f"{what}: shape {tuple(actual.shape)} != {tuple(expected.shape)}"
    assert torch.isfinite(actual).all(), f"{what}: non-finite values in the CUDA result"
    err = (actual - expected).abs()
    ok = (err <= atol) | (err <= rtol * expected.abs())
    if not ok.all():
        bad = (~ok).sum().item()
        idx = torch.nonzero(~ok)[0].tolist()
        raise AssertionError(
            f"{what}: {bad}/{ok.numel()} elements outside abs {atol} / rel {rtol}; max abs err {err.max().item():.3e}; "
            f"first bad at {idx}: got {actual[tuple(idx)].item():.6f} want {expected[tuple(idx)].item():.6f}")
    return err.max().item()


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


BASE_HPARAMS = dict(n_frames_total=1, n_frames_now=1, person_inputs=["agnostic", "densepose"], cloth_inputs=["cloth"],
                    ngf=64, self_attn=True, num_attn=2, flow_warp=False, activation="gelu", is_train=False,
                    grid_size=5, fine_height=256, fine_width=192)


def make_hparams(**over):
    import argparse

    d = dict(BASE_HPARAMS)
    d.update(over)
    return argparse.Namespace(**d)


def build_model(kind, seed=420, **over):
    """Our nn.Module mirror with the deterministic synthetic weights (oracle/weights.py) loaded strictly."""
    from oracle import weights
    from shineon_virtual_tryon_b200.models import find_model_using_name

    if kind == "warp":
        over.setdefault("person_inputs", ["agnostic", "cocopose"])
    m = find_model_using_name(kind)(make_hparams(**over))
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = weights.synth_state_dict(shapes, seed)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd
