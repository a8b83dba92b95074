"""Data-parallel training driver of the U-Net stage (reference: Lightning's DDP loop around
UnetMaskModel.training_step + Adam, train.py:52-62, models/base_model.py:165-184).

One process per GPU.  Parameters live in ONE flat f32 buffer and their gradients in another (every `.grad` is a view),
so the gradient exchange is a handful of bucketed NCCL all-reduces over NVLink/NVSwitch — launched from inside the
backward pass as soon as a bucket's last gradient is final, i.e. overlapped with the remaining backward kernels — and
the optimiser is one fused Adam launch over the whole model.  `accumulated_batches` micro-batches accumulate into the
same gradient buffer; the exchange happens on the last one only.
"""
import torch

from . import ops
from .distributed import FlatGradAllReducer
from .networks import _engine_util
from .networks.cpvton import unet as unet_mod


def backward_order(unet_generator):
    """Parameters of a UnetGenerator in the order the hand-written backward finalises their gradients:
    up-path of the outermost block first, then inwards; the down-paths on the way back out."""
    blocks, b = [], unet_generator.model
    while b is not None:
        blocks.append(b)
        b = b._parts["sub"]
    order = []
    for b in blocks:
        order += b.up_params()
    for b in reversed(blocks):
        order += b.down_params()
    return order


class Trainer:
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, accumulated_batches=1, bucket_bytes=32 << 20,
                 lr_lambda=None, cuda_graph=False, graph_warmup=2, graph_allreduce=False):
        """cuda_graph=True: after `graph_warmup` eager micro-batches the whole training step (forward, losses, backward:
        ~450 launches, launch-bound at the recipe's batch 4) is captured once into a CUDA graph and replayed; the batch
        is copied into static buffers and the gradient exchange runs after the replay (bucketed, asynchronous, not overlapped
        with the backward: 90.6 MB over NVLink is ~0.4 ms of a 6 ms step).  graph_allreduce=True (EXPERIMENTAL, off by default)
        captures the bucketed NCCL all-reduces issued from inside the backward as side-stream nodes of the same graph; on this
        image (torch 2.11 / NCCL 2.28.9) the 2-GPU capture dead-locked, so the measured configuration is the default."""
        self.cuda_graph, self.graph_warmup, self.graph_allreduce = bool(cuda_graph), int(graph_warmup), bool(graph_allreduce)
        # two captured variants: [True] the first micro-batch after an optimiser step (the 16-bit weight re-pack is part of
        # the captured work) and [False] the following micro-batches of an accumulation window (weights unchanged: they
        # read the buffers the [True] graph re-packs into)
        self._graphs, self._static_batch = {}, None
        self.model = model
        if self.cuda_graph:
            from . import _lib

            fmt, _ = ops.resolve_precision(model.unet.precision if model.unet.precision is not None else "bf16x3")
            if fmt == _lib.FMT_FP16:
                raise NotImplementedError("Trainer(cuda_graph=True) needs a bf16 plane format: the fp16 modes derive a "
                                          "power-of-two weight scale on the host at every re-pack, which a graph would freeze")
        ordered = [p for p in backward_order(model.unet) if p.requires_grad]
        seen = {id(p) for p in ordered}
        rest = [p for p in model.parameters() if p.requires_grad and id(p) not in seen]
        self.params = ordered + rest
        self.reducer = FlatGradAllReducer(self.params, bucket_bytes=bucket_bytes)
        dev = self.params[0].device
        self.flat_param = torch.empty(self.reducer.numel, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for i, p in enumerate(self.params):
                n = p.numel()
                view = self.flat_param[off:off + n].view_as(p)
                view.copy_(p.data)
                p.data = view              # the parameter now aliases the flat buffer
                p.grad = self.reducer.grad_view(i)
                off += n
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.accumulated_batches = max(1, int(accumulated_batches))
        self.lr_lambda = lr_lambda
        self.micro, self.steps, self.epoch = 0, 0, 0
        _engine_util.bump_weights_epoch()

    def named_params(self, model=None):
        """(name, parameter) in flat-buffer order (checkpoints store the Adam moments in this order)."""
        names = {id(p): n for n, p in (model or self.model).named_parameters()}
        return [(names.get(id(p), f"param{i}"), p) for i, p in enumerate(self.params)]

    def zero_grad(self):
        self.reducer.flat.zero_()

    def train_batch(self, batch, batch_idx=0):
        """One micro-batch: forward + losses + backward (gradients accumulate); on every `accumulated_batches`-th call
        the gradient all-reduce (overlapped with that backward) and the fused Adam step.  Returns the step's result."""
        last = (self.micro + 1) % self.accumulated_batches == 0
        first = self.micro % self.accumulated_batches == 0  # weights changed since the previous micro-batch
        # the first capture must be of a `first` micro-batch, so the packed-weight buffers every later graph reads are
        # the ones that graph rewrites on each replay
        if self.cuda_graph and self.micro >= self.graph_warmup and (first or any(k[0] for k in self._graphs)):
            in_graph = last and self.graph_allreduce and self.reducer.active()
            res = self._replay(batch, batch_idx, first, in_graph)
            self.micro += 1
            if last:
                if in_graph:
                    inv_world = self.reducer.inv_world()
                else:
                    self.reducer.start()
                    inv_world = self.reducer.finish()
                self.optimizer_step(inv_world / self.accumulated_batches)
                self.zero_grad()
            return res
        if last:
            self.reducer.begin_overlap()
            unet_mod.GRAD_READY_HOOK = self.reducer.mark_ready
        try:
            res = self.model.training_step(batch, batch_idx)
        finally:
            unet_mod.GRAD_READY_HOOK = None
        self.micro += 1
        if last:
            inv_world = self.reducer.end_overlap()
            self.optimizer_step(inv_world / self.accumulated_batches)
            self.zero_grad()
        return res

    def _replay(self, batch, batch_idx, first, exchange=False):
        """exchange: this micro-batch ends an accumulation window on a multi-rank job -- its graph also holds the bucketed
        gradient all-reduces, issued by the backward as the buckets complete."""
        tensors = {k: v for k, v in batch.items() if isinstance(v, torch.Tensor)}
        if self._static_batch is None:
            self._static_batch = {k: v.clone() for k, v in tensors.items()}
        ent = self._graphs.get((first, exchange))
        if ent is None:
            from . import _lib

            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            epoch0 = _engine_util.WEIGHTS_EPOCH[0]
            n0 = _lib.launch_count()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):  # the NCCL watchdog thread may poll events
                if exchange:
                    self.reducer.begin_overlap()
                    unet_mod.GRAD_READY_HOOK = self.reducer.mark_ready
                try:
                    res = self.model.training_step(self._static_batch, batch_idx)
                finally:
                    unet_mod.GRAD_READY_HOOK = None
                if exchange:
                    self.reducer.end_overlap()  # the waits join NCCL's stream back into the captured one
            ent = self._graphs[(first, exchange)] = (graph, res, _lib.launch_count() - n0)
            assert _engine_util.WEIGHTS_EPOCH[0] == epoch0
            with_repack = [v[2] for k, v in self._graphs.items() if k[0]]
            if first:
                # the re-pack of every trainable layer must have been recorded: compare with the no-repack variant later
                self.repack_launches = ent[2]
            elif with_repack:
                assert ent[2] < max(with_repack), "the weight re-pack was not captured in the post-step graph"
            # the capture only recorded the work: the replay below runs it for this batch
        for k, v in tensors.items():
            self._static_batch[k].copy_(v, non_blocking=True)
        ent[0].replay()
        return ent[1]

    def optimizer_step(self, grad_scale=1.0):
        self.steps += 1
        lr = self.lr * (self.lr_lambda(self.epoch) if self.lr_lambda else 1.0)
        ops.adam_step(self.flat_param, self.reducer.flat, self.exp_avg, self.exp_avg_sq, self.steps, lr=lr, betas=self.betas,
                      eps=self.eps, grad_scale=grad_scale)
        _engine_util.bump_weights_epoch()   # packed 16-bit weight copies are stale now
