#!/usr/bin/env python
"""train.py — training entry point with the reference's command line (train.py:32-145):

    python train.py --model unet --name NAME --self_attn --activation gelu -b 4 [--accumulated_batches 16] ...
    python -m torch.distributed.run --nproc-per-node 8 train.py ...      (one process per GPU, NCCL)

Parses TrainOptions and builds the model exactly like the reference does.  The optimisation loop replaces Lightning's:
`UnetMaskModel.training_step` runs forward, losses and the whole backward on the hand-written kernels;
`training.Trainer` keeps parameters/gradients in flat buffers, all-reduces the gradients over NCCL (overlapped with the
backward) and applies the fused Adam with the reference's linear-decay schedule (models/base_model.py:165-184).
The built-in dataset is synthetic (the reference's VVT/VITON readers are its CPU data path); any
dataset yielding the same batch dict can be passed to `fit`.  Only the U-Net stage trains natively (rows U6/U7)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def fit(model, opt, dataset, device, max_steps=None, log_every=10):
    import torch
    from torch.utils.data import DataLoader
    from torch.utils.data.distributed import DistributedSampler

    from shineon_virtual_tryon_b200 import distributed
    from shineon_virtual_tryon_b200.training import Trainer

    rank, world = distributed.init_process_group(device=device)
    sampler = DistributedSampler(dataset, num_replicas=world, rank=rank, shuffle=not opt.no_shuffle, seed=420) if world > 1 else None
    loader = DataLoader(dataset, batch_size=opt.batch_size, shuffle=(sampler is None and not opt.no_shuffle),
                        sampler=sampler, num_workers=0, drop_last=True)

    def lr_lambda(epoch):  # BaseModel._make_step_scheduler (base_model.py:170-184)
        return 1.0 - max(0, epoch - opt.keep_epochs) / float(opt.decay_epochs + 1)

    trainer = Trainer(model, lr=opt.lr, accumulated_batches=opt.accumulated_batches, lr_lambda=lr_lambda)
    epochs = opt.keep_epochs + opt.decay_epochs
    t0, seen = time.time(), 0
    for epoch in range(epochs):
        trainer.epoch = epoch
        if sampler is not None:
            sampler.set_epoch(epoch)
        for bi, batch in enumerate(loader):
            batch = {k: (v.to(device, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
            res = trainer.train_batch(batch, bi)
            seen += opt.batch_size * world
            if rank == 0 and trainer.micro % log_every == 0:
                logs = " ".join(f"{k.split('/')[-1]}={float(v):.4f}" for k, v in res["log"].items())
                print(f"epoch {epoch} step {trainer.steps} ({seen / (time.time() - t0):.1f} samples/s): {logs}", flush=True)
            if opt.fast_dev_run or (max_steps is not None and trainer.steps >= max_steps):
                return trainer
    return trainer


def main(argv=None, max_steps=None):
    import torch

    from shineon_virtual_tryon_b200.models import find_model_using_name
    from shineon_virtual_tryon_b200.options import TrainOptions
    from shineon_virtual_tryon_b200.synthetic_data import SyntheticTryonDataset

    torch.manual_seed(420)  # train.py:27-29
    opt = TrainOptions().parse(argv)
    model_class = find_model_using_name(opt.model)
    if opt.model != "unet_mask":
        raise SystemExit(f"native training covers the U-Net try-on stage (--model unet); `{opt.model}` trains only in the "
                         "reference (SURVEY.md section 8a rows U6/U7)")
    model = model_class(opt)
    if getattr(opt, "checkpoint", None):
        state = torch.load(opt.checkpoint, map_location="cpu")
        model.load_state_dict(state.get("state_dict", state), strict=True)
    local = int(os.environ.get("LOCAL_RANK", opt.gpu_ids[0] if opt.gpu_ids else 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    model = model.to(dev).train()
    model.set_train_precision(getattr(opt, "b200_train_precision", "bf16"))
    n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
    if int(os.environ.get("RANK", 0)) == 0:
        print(f"built {type(model).__name__} ({n_params / 1e6:.2f} M trainable parameters); Adam(lr={opt.lr}), linear decay "
              f"after {opt.keep_epochs} epochs, accumulate {opt.accumulated_batches}")
    trainer = fit(model, opt, SyntheticTryonDataset(opt), dev, max_steps=max_steps if max_steps is not None else opt.max_steps)
    if opt.name and getattr(opt, "save_final", True) and int(os.environ.get("RANK", 0)) == 0:
        ckpt_dir = os.path.join(getattr(opt, "experiments_dir", "experiments"), opt.name)
        os.makedirs(ckpt_dir, exist_ok=True)
        torch.save({"state_dict": {k: v.detach().cpu().clone() for k, v in model.state_dict().items()},
                    "global_step": trainer.steps}, os.path.join(ckpt_dir, "final.ckpt"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
