"""tests/golden/sams_*.npz: the UNMODIFIED reference SamsGenerator (models/networks/sams/sams_generator.py, imported
read-only through oracle/ref_shim.py) on the seeded cases of oracle/cases.py.  Build-container only.

    python -m oracle.make_golden_sams
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cases, ref_shim, weights  # noqa: E402
from oracle.make_golden import SEED, _save  # noqa: E402


def build_reference(name):
    ref_shim.install()
    from models.networks.sams.sams_generator import SamsGenerator

    over = cases.SAMS_CASES[name][0]
    g = SamsGenerator(ref_shim.hparams(**over)).eval()
    shapes = weights.shapes_of(g)
    g.load_state_dict(weights.fix_spectral(weights.synth_state_dict(shapes, SEED)), strict=True)
    return g, shapes


def main():
    torch.manual_seed(SEED)
    with torch.no_grad():
        for name in cases.SAMS_CASES:
            g, shapes = build_reference(name)
            prev, prev_maps, maps = cases.sams_inputs(name)
            out = g(prev, prev_maps, maps)
            H = cases.SAMS_CASES[name][2]
            _save(name, shapes, out=cases.subsample(out, 4 if H >= 256 else 1))


def model_golden(name="sams_small"):
    """SamsModel.generate_n_frames of the unmodified reference (Resample2d routed to the pinned CPU oracle, its only
    implementation being CUDA): the whole frame buffer after n_frames_total generator passes with flow blending."""
    ref_shim.install()
    ref_shim.patch_native_ops_with_oracle()
    from models.sams_model import SamsModel

    over = cases.SAMS_CASES[name][0]
    with torch.no_grad():
        m = SamsModel(ref_shim.hparams(**over)).eval()
        shapes = weights.shapes_of(m)
        m.load_state_dict(weights.fix_spectral(weights.synth_state_dict(shapes, SEED)), strict=True)
        last, _, frames = m.generate_n_frames(cases.sams_model_batch(name))
        _save(name + "_model", shapes, last=last, frames=frames)


if __name__ == "__main__":
    model_golden()
    main()
