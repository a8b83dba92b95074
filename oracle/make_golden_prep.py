"""Generates tests/golden/frame_prep.npz by calling the UNMODIFIED reference dataset methods (imported read-only through
oracle/ref_shim.py) on the seeded synthetic frames of oracle/frame_prep.synth_frame.  Build-container only.

    python -m oracle.make_golden_prep

Reference entry points exercised (datasets/tryon_dataset.py): to_tensor_and_norm_rgb (:110-112), get_input_cloth_mask
(:168-175), get_person_head (:323-344), get_person_body_silhouette (:346-367), convert_pose_data_to_pose_map_and_vis
(:389-447); flownet2_pytorch/utils/flow_utils.py readFlow (:7-26) + flow_norm (tryon_dataset.py:121,288-289).
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import frame_prep as fp, ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "frame_prep.npz")
SEEDS = (1, 2)
H, W = 256, 192


def reference_self():
    from torchvision import transforms

    s = types.SimpleNamespace(fine_height=H, fine_width=W, radius=5, cloth_mask_threshold=240)
    s.center_crop = transforms.CenterCrop((H, W))
    s.rgb_norm = transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))
    s.to_tensor_and_norm_rgb = transforms.Compose([s.center_crop, transforms.ToTensor(), s.rgb_norm])
    s.to_tensor_and_norm_gray = transforms.Compose([s.center_crop, transforms.ToTensor(), transforms.Normalize([0.5], [0.5])])
    s.flow_norm = transforms.Normalize((0.5, 0.5), (0.5, 0.5))
    return s


def main():
    ref_shim.install()
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    from PIL import Image

    from datasets.tryon_dataset import TryonDataset
    from models.flownet2_pytorch.utils.flow_utils import readFlow

    s = reference_self()
    out = {}
    for seed in SEEDS:
        image, parse, cloth, densepose, pose = fp.synth_frame(seed, H, W)
        im = s.to_tensor_and_norm_rgb(Image.fromarray(image))
        cl = s.to_tensor_and_norm_rgb(Image.fromarray(cloth))
        pose_map, vis = TryonDataset.convert_pose_data_to_pose_map_and_vis(s, pose)
        out[f"s{seed}_image_sub"] = im[:, ::4, ::4].numpy()
        out[f"s{seed}_cloth_sub"] = cl[:, ::4, ::4].numpy()
        out[f"s{seed}_cloth_mask"] = TryonDataset.get_input_cloth_mask(s, cl).numpy().astype(np.uint8)
        out[f"s{seed}_im_head_sub"] = TryonDataset.get_person_head(s, im, parse)[:, ::4, ::4].numpy()
        out[f"s{seed}_silhouette"] = TryonDataset.get_person_body_silhouette(s, parse).numpy()
        out[f"s{seed}_pose_map_minmax"] = np.array([pose_map.min().item(), pose_map.max().item()], np.float32)
        out[f"s{seed}_pose_map_shape"] = np.array(pose_map.shape, np.int32)
        out[f"s{seed}_im_cocopose"] = np.packbits(vis.numpy()[0] > 0)
    # .flo round trip through the reference reader
    r = np.random.RandomState(7)
    flow = (r.randn(24, 20, 2) * 3).astype(np.float32)
    buf = np.float32(202021.25).tobytes() + np.int32(20).tobytes() + np.int32(24).tobytes() + flow.tobytes()
    with tempfile.NamedTemporaryFile(suffix=".flo") as f:
        f.write(buf)
        f.flush()
        ref = s.flow_norm(torch.from_numpy(readFlow(f.name)).permute(2, 0, 1))
    out["flo_bytes"] = np.frombuffer(buf, np.uint8)
    out["flo_decoded"] = ref.numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
