"""Generates tests/golden/image_u8.npz with the UNMODIFIED reference writer: visualization.save_images
(visualization.py:59-88) writes PNGs into a temp dir, the pixel arrays are read back with PIL.  Build-container only.

    python -m oracle.make_golden_image
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import image_io, ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "image_u8.npz")
SEEDS = (3, 4)


def main():
    ref_shim.install()
    import visualization as vis
    from PIL import Image

    arrs = {}
    for seed in SEEDS:
        for C in (3, 1):
            x = image_io.synth_images(seed, C=C)
            d = tempfile.mkdtemp()
            names = [f"{i}.png" for i in range(x.shape[0])]
            vis.save_images(torch.from_numpy(x), names, [d] * len(names))
            got = np.stack([np.array(Image.open(os.path.join(d, n))) for n in names])
            arrs[f"s{seed}_c{C}"] = got if C == 3 else got[..., None]
    np.savez_compressed(OUT, **arrs)
    print("wrote", OUT, {k: v.shape for k, v in arrs.items()})


if __name__ == "__main__":
    main()
