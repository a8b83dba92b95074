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
