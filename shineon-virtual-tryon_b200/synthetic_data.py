"""SyntheticTryonDataset — random tensors with the reference's batch-dict contract (datasets/tryon_dataset.py:47-61,
481-537): keys image, prev_image, cloth, cloth_mask, agnostic, cocopose, densepose, flow, im_cloth, grid_vis, each
[n_frames, C, H, W] per sample.  Stands in for the reference's CPU data path (VVT / VITON / MPV readers), which is out of
scope; any object yielding the same dict can be passed to test.py / the models instead."""
import torch
from torch.utils.data import Dataset

from .models.base_model import CHANNELS


class SyntheticTryonDataset(Dataset):
    KEYS = dict(image="RGB", prev_image="RGB", cloth="CLOTH", im_cloth="RGB", cloth_mask="CLOTH_MASK", agnostic="AGNOSTIC",
                cocopose="COCOPOSE", densepose="DENSEPOSE", flow="FLOW", grid_vis="RGB")

    def __init__(self, opt, length=None, seed=420):
        self.h, self.w = opt.fine_height, opt.fine_width
        self.n = opt.n_frames_total
        self.length = length if length is not None else getattr(opt, "synthetic_samples", 16)
        self.seed = seed

    def __len__(self):
        return self.length

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 100003 + i)
        item = {}
        for key, kind in self.KEYS.items():
            c = CHANNELS[kind]
            if key == "cloth_mask":
                t = (torch.rand(self.n, c, self.h, self.w, generator=g) > 0.5).float()
            elif key == "flow":
                t = torch.randn(self.n, c, self.h, self.w, generator=g) * 3
            else:
                t = torch.rand(self.n, c, self.h, self.w, generator=g) * 2 - 1
            item[key] = t
        item["image_name"] = [f"synthetic_{i:06d}_f{f}.png" for f in range(self.n)]
        item["cloth_name"] = f"synthetic_cloth_{i:06d}.png"
        item["dataset_name"] = "synthetic"
        return item
