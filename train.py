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


def save_checkpoint(path, model, trainer, opt, epoch):
    """Lightning-style checkpoint the reference's `load_from_checkpoint` understands (state_dict + hyper_parameters +
    global_step / epoch) plus what a resume needs and Lightning would store under `optimizer_states`: Adam moments and
    step count of the flat buffers, in parameter order."""
    import torch

    hp = {k: v for k, v in vars(opt).items() if isinstance(v, (int, float, str, bool, list, tuple, type(None)))}
    torch.save({"state_dict": {k: v.detach().cpu().clone() for k, v in model.state_dict().items()},
                "hyper_parameters": hp, "global_step": trainer.steps, "epoch": epoch,
                "b200_optimizer": {"exp_avg": trainer.exp_avg.cpu(), "exp_avg_sq": trainer.exp_avg_sq.cpu(),
                                   "steps": trainer.steps, "micro": trainer.micro,
                                   "param_names": [n for n, _ in trainer.named_params(model)]}}, path)


def fit(model, opt, dataset, device, max_steps=None, log_every=10, resume=None, ckpt_dir=None):
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
    fit.last_trainer = trainer
    epochs = opt.keep_epochs + opt.decay_epochs
    start_epoch = 0
    if resume is not None and "b200_optimizer" in resume:  # continue bias correction and the LR schedule where they stopped
        st = resume["b200_optimizer"]
        if st["param_names"] == [n for n, _ in trainer.named_params(model)]:
            trainer.exp_avg.copy_(st["exp_avg"])
            trainer.exp_avg_sq.copy_(st["exp_avg_sq"])
            trainer.steps, start_epoch = int(st["steps"]), int(resume.get("epoch", 0))
        elif rank == 0:
            print("checkpoint optimizer state does not match this model's parameters: Adam restarts", flush=True)
    t0, seen = time.time(), 0
    save_every = max(1, int(getattr(opt, "save_count", 10000)))
    for epoch in range(start_epoch, epochs):
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
            if ckpt_dir and rank == 0 and trainer.micro % trainer.accumulated_batches == 0 and trainer.steps % save_every == 0:
                save_checkpoint(os.path.join(ckpt_dir, f"step_{trainer.steps:07d}.ckpt"), model, trainer, opt, epoch)
            if opt.fast_dev_run or (max_steps is not None and trainer.steps >= max_steps):
                return trainer
    return trainer


def main(argv=None, max_steps=None, dataset=None):
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
    if getattr(opt, "dataset", "synthetic") != "synthetic" and dataset is None:
        raise SystemExit(f"--dataset {opt.dataset}: the reference's VVT / VITON / MPV readers are its CPU data path and are not "
                         "part of this build (SURVEY.md section 2, out of scope).  Pass a dataset object yielding the reference's "
                         "batch dict to train.main(dataset=...), or use --dataset synthetic.")
    model = model_class(opt)
    state = None
    if getattr(opt, "checkpoint", None):
        state = torch.load(opt.checkpoint, map_location="cpu")
        model.load_state_dict(state.get("state_dict", state), strict=True)
    if getattr(opt, "vgg_weights", None):
        model.criterionVGG.load_pretrained(opt.vgg_weights)
    elif not model.criterionVGG.pretrained_loaded:
        model.criterionVGG.load_pretrained(None)  # torchvision's cached ImageNet weights if this machine has them
    local = int(os.environ.get("LOCAL_RANK", opt.gpu_ids[0] if opt.gpu_ids else 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    model = model.to(dev).train()
    model.set_train_precision(getattr(opt, "b200_train_precision", "bf16"))
    n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
    if int(os.environ.get("RANK", 0)) == 0:
        print(f"built {type(model).__name__} ({n_params / 1e6:.2f} M trainable parameters); Adam(lr={opt.lr}), linear decay "
              f"after {opt.keep_epochs} epochs, accumulate {opt.accumulated_batches}")
    is_rank0 = int(os.environ.get("RANK", 0)) == 0
    ckpt_dir = None
    if opt.name and getattr(opt, "save_final", True):
        ckpt_dir = os.path.join(getattr(opt, "experiments_dir", "experiments"), opt.name)
        if is_rank0:
            os.makedirs(ckpt_dir, exist_ok=True)
    holder = {}
    try:
        trainer = fit(model, opt, dataset if dataset is not None else SyntheticTryonDataset(opt), dev,
                      max_steps=max_steps if max_steps is not None else opt.max_steps, resume=state, ckpt_dir=ckpt_dir)
        holder["t"] = trainer
    except KeyboardInterrupt:  # the reference saves on interrupt too (train.py:117-127)
        trainer = getattr(fit, "last_trainer", None)
        if trainer is not None and ckpt_dir and is_rank0:
            save_checkpoint(os.path.join(ckpt_dir, "interrupted.ckpt"), model, trainer, opt, trainer.epoch)
        raise
    if ckpt_dir and is_rank0:
        save_checkpoint(os.path.join(ckpt_dir, "final.ckpt"), model, trainer, opt, trainer.epoch)
    return 0


if __name__ == "__main__":
    sys.exit(main())
