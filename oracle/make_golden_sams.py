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


if __name__ == "__main__":
    main()
