"""Multi-GPU plumbing of the inference path: one process per GPU (torch.distributed), clips sharded over ranks.

Frames / clips are independent on the try-on inference path, so there is NO data-path collective:
each rank owns `shard_clips(...)` of the clip list.  The process group is only used for the barrier and for the
max-over-ranks reduction of device-timed intervals (NCCL on GPUs, gloo in the CPU tests).
"""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init_process_group(backend=None, device=None):
    """Initialise from the torchrun environment (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).  No-op for world 1."""
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def shard_clips(n_clips, rank, world):
    """Round-robin assignment of clip indices to ranks (the reference's DistributedSampler without shuffling,
    models/base_model.py:113-115, minus its padding: every clip is processed exactly once)."""
    return list(range(rank, n_clips, world))


def barrier():
    if dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(value, device="cpu"):
    """Max of a python float over all ranks (device-timed milliseconds -> slowest rank)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


class FlatGradAllReducer:
    """Data-parallel gradient exchange of the training path: ONE flat f32 gradient buffer,
    summed over ranks with bucketed asynchronous all-reduces (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU
    tests).  The 1/world_size of the mean is folded into the optimiser's `grad_scale` (ops.adam_step), so the
    collective moves raw sums.  Buckets are issued last-layer-first so the exchange of the layers whose gradients
    are ready first overlaps the rest of the backward pass."""

    def __init__(self, params, bucket_bytes=32 << 20):
        """`params` in the order their gradients become final during the backward pass (see
        training.backward_order): bucket k can then be launched as soon as the last of its parameters is done."""
        self.params = [p for p in params]
        self.numel = sum(p.numel() for p in self.params)
        first = self.params[0]
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=first.device)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        per = max(1, bucket_bytes // 4)
        self.buckets = [(s, min(s + per, self.numel)) for s in range(0, self.numel, per)]
        self._work = []
        # overlap bookkeeping: parameter i ends at offset ends[i]; a bucket is complete once every parameter that
        # overlaps it has been reported by mark_ready (parameters are reported in buffer order)
        self._ends, off = [], 0
        for p in self.params:
            off += p.numel()
            self._ends.append(off)
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._ready_upto = 0   # elements [0, _ready_upto) hold final gradients
        self._next_bucket = 0
        self._overlap = False

    # ---- overlapped mode: all-reduce of bucket k starts while the backward pass is still producing later buckets
    def begin_overlap(self):
        self._work, self._ready_upto, self._next_bucket, self._overlap = [], 0, 0, True
        self._done = [False] * len(self.params)

    def mark_ready(self, params):
        """Called by the backward pass when the gradients of `params` are final."""
        if not self._overlap:
            return
        for p in params:
            i = self._index.get(id(p))
            if i is not None:
                self._done[i] = True
        i = 0
        while self._ready_upto < self.numel:
            # advance over the prefix of finished parameters
            while i < len(self.params) and self._ends[i] <= self._ready_upto:
                i += 1
            if i >= len(self.params) or not self._done[i]:
                break
            self._ready_upto = self._ends[i]
        active = dist.is_initialized() and dist.get_world_size() > 1
        while self._next_bucket < len(self.buckets) and self.buckets[self._next_bucket][1] <= self._ready_upto:
            s, e = self.buckets[self._next_bucket]
            if active:
                self._work.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, async_op=True))
            self._next_bucket += 1

    def end_overlap(self):
        """Launches whatever mark_ready has not covered (e.g. frozen parameters at the tail) and waits."""
        active = dist.is_initialized() and dist.get_world_size() > 1
        while self._next_bucket < len(self.buckets):
            s, e = self.buckets[self._next_bucket]
            if active:
                self._work.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, async_op=True))
            self._next_bucket += 1
        self._overlap = False
        return self.finish()

    @staticmethod
    def active():
        return dist.is_initialized() and dist.get_world_size() > 1

    @staticmethod
    def inv_world():
        return 1.0 / (dist.get_world_size() if dist.is_initialized() else 1)

    def grad_view(self, i):
        """Where the backward pass writes the gradient of parameter i (no per-parameter .grad tensors to pack)."""
        return self.views[i]

    def start(self):
        """Launch the bucketed sums (async), last bucket first."""
        self._work = []
        if dist.is_initialized() and dist.get_world_size() > 1:
            for s, e in reversed(self.buckets):
                self._work.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, async_op=True))

    def finish(self):
        for w in self._work:
            w.wait()
        self._work = []
        return 1.0 / (dist.get_world_size() if dist.is_initialized() else 1)  # grad_scale for the optimiser
