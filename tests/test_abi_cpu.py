"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol that
include/shineon_b200.h declares; the ctypes table covers them all; host-side argument validation fails loudly."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "shineon_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(shineon_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from shineon_virtual_tryon_b200 import _lib

    lib = _lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/shineon_b200.h but not exported by libshineon_b200.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype in _lib.SIGNATURES"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_channel():
    from shineon_virtual_tryon_b200 import _lib

    lib = _lib.load()
    assert lib.shineon_version() >= 100
    assert isinstance(_lib.launch_count(), int)
    # argument errors are reported without touching a device
    rc = lib.shineon_channelnorm_fwd(None, None, 1, 3, 4, 4, 2, None)
    assert rc == -1 and b"null pointer" in lib.shineon_last_error()
    import ctypes

    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.shineon_correlation_out_shape(256, 32, 24, 20, 1, 20, 1, 2, ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)) == 0
    assert (oc.value, oh.value, ow.value) == (441, 32, 24)  # FlowNetC.py:31 on the 256x192 path


def test_ops_refuse_cpu_tensors():
    import torch

    from shineon_virtual_tryon_b200 import _lib, ops

    with pytest.raises(_lib.ShineonError):
        ops.resample2d_fwd(torch.zeros(1, 3, 4, 4), torch.zeros(1, 2, 4, 4))
    with pytest.raises(_lib.ShineonError):
        ops.nchw_to_planes(torch.zeros(1, 3, 4, 4))


def test_state_dict_keys_match_reference_golden():
    """Module mirrors expose exactly the reference's state_dict keys / shapes (taken from the reference by
    oracle/make_golden.py and stored in the fixtures)."""
    import argparse

    from oracle import cases
    from shineon_virtual_tryon_b200.models import find_model_using_name
    from shineon_virtual_tryon_b200.networks.flownet2.nets import FlowNet2
    from tests.golden_util import load_golden
    from tests.util import make_hparams

    for name, (over, _) in cases.TOM_CASES.items():
        m = find_model_using_name("unet_mask")(make_hparams(**over))
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == load_golden(name)[1], name
    w = find_model_using_name("warp")(make_hparams(person_inputs=["agnostic", "cocopose"]))
    assert {k: tuple(v.shape) for k, v in w.state_dict().items()} == load_golden("gmm_b2")[1]
    assert {k: tuple(v.shape) for k, v in FlowNet2().state_dict().items()} == load_golden("flownet2")[1]
