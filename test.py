#!/usr/bin/env python
"""test.py — inference entry point with the reference's command line (test.py:10, train.py:32-73 with train=False):

    python test.py --model warp|unet_mask --name NAME [--checkpoint CKPT] [-b 4] [--self_attn --activation gelu] ...

Builds the options (options/*), the model class by name, loads the checkpoint's state_dict if given (reference
checkpoints load unchanged), and runs `test_step` over the dataset.  The built-in dataset is synthetic (the reference's
VVT/VITON readers are its CPU data path); results are written as PNGs like visualization.save_images does
((x+1)*127.5 -> uint8) when --result_dir is set and PIL is available."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def save_images(img_tensors, img_names, save_dir):
    """visualization.save_images (visualization.py:60-88)."""
    import numpy as np
    from PIL import Image

    os.makedirs(save_dir, exist_ok=True)
    for img_tensor, img_name in zip(img_tensors, img_names):
        tensor = (img_tensor.detach().float().cpu().clamp(-1, 1) + 1) * 0.5 * 255
        array = tensor.numpy().astype("uint8")
        if array.shape[0] == 1:
            array = array.squeeze(0)
        elif array.shape[0] == 3:
            array = array.swapaxes(0, 1).swapaxes(1, 2)
        Image.fromarray(array).save(os.path.join(save_dir, img_name))


def main(argv=None):
    import torch
    from torch.utils.data import DataLoader

    from shineon_virtual_tryon_b200.models import find_model_using_name
    from shineon_virtual_tryon_b200.options import TestOptions
    from shineon_virtual_tryon_b200.synthetic_data import SyntheticTryonDataset

    opt = TestOptions().parse(argv)
    model = find_model_using_name(opt.model)(opt)
    if opt.checkpoint:
        state = torch.load(opt.checkpoint, map_location="cpu")
        model.load_state_dict(state.get("state_dict", state), strict=True)
    model.override_hparams(opt) if opt.checkpoint else None
    dev = torch.device("cuda", opt.gpu_ids[0] if opt.gpu_ids else 0)
    model = model.to(dev).eval()
    model.set_precision(opt.b200_precision)
    loader = DataLoader(SyntheticTryonDataset(opt), batch_size=opt.batch_size, num_workers=0)
    n, t0 = 0, time.time()
    with torch.no_grad():
        for bi, batch in enumerate(loader):
            batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
            out = model.test_step(batch, bi)
            key = "p_tryon" if "p_tryon" in out else "warped_cloth"
            n += out[key].shape[0]
            if opt.result_dir:
                names = [f"{opt.model}_{bi:04d}_{i:02d}.png" for i in range(out[key].shape[0])]
                save_images(out[key], names, os.path.join(opt.result_dir, opt.name, opt.datamode))
            if opt.fast_dev_run:
                break
    torch.cuda.synchronize()
    print(f"{opt.model}: {n} samples in {time.time() - t0:.2f} s")
    return 0


if __name__ == "__main__":
    sys.exit(main())
