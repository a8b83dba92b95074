"""In-memory two-stage try-on: GMM -> TPS warp -> TOM, frames batched along N.

The reference runs the two stages as separate `test.py` invocations and passes the warped cloth through 8-bit
PNGs on disk (docs/2_inference.md:13-39, models/warp_model.py:146-148, datasets/vvt_dataset.py:133-150).
This class chains the same two forwards (WarpModel.forward + grid_sample(border), UnetMaskModel.forward)
on the device; `run_host` is the host-buffer entry point used for the end-to-end number (pinned host
tensors in, pinned host tensor out, copies on the caller's stream).
"""
import torch


class TryOnPipeline:
    def __init__(self, warp_model, tom_model, cuda_graph=False):
        """cuda_graph=True: the ~180 launches of a step are captured once per distinct set of input buffers and
        replayed (the small launches at the bottom of the U-Net are otherwise issued slower than the GPU runs them).
        Replayed calls return the graph's static output tensors: they are overwritten by the next call with the same
        input buffers."""
        self.warp_model = warp_model
        self.tom_model = tom_model
        self._host_state = None
        self.cuda_graph = bool(cuda_graph)
        self._graphs = {}
        self.replayed_launches = 0  # kernel launches executed through graph replays (not seen by shineon_launch_count)

    def set_precision(self, precision):
        self.warp_model.set_precision(precision)
        self.tom_model.set_precision(precision)
        self._graphs.clear()  # captured graphs hold the previous mode's kernels

    def _stages(self, person_gmm, cloth, person_tom):
        warped_cloth, _, _, _ = self.warp_model.warp(person_gmm, cloth, cloth)
        _, tryon_masks, p_tryons, _ = self.tom_model(person_tom, warped_cloth)
        return p_tryons, tryon_masks, warped_cloth

    def _graph_call(self, tag, fn, args):
        """fn(*args) through a CUDA graph keyed by (tag, argument buffers).  Eager while ops.PROFILE is collecting."""
        from . import _lib, ops

        if not self.cuda_graph or ops.PROFILE is not None:
            return fn(*args)
        key = (tag,) + tuple((a.data_ptr(), tuple(a.shape), a.dtype) for a in args)
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= 8:
                self._graphs.clear()
            for _ in range(2):  # weight packing, allocator warm-up: nothing of that may happen inside the capture
                fn(*args)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            try:
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):  # other threads (NCCL watchdog) may poll events
                    out = fn(*args)
            except RuntimeError as e:  # a capture the runtime refuses must not take the step down: launch eagerly instead
                import warnings

                warnings.warn(f"TryOnPipeline: CUDA-graph capture failed ({str(e).splitlines()[0]}); running eagerly")
                torch.cuda.synchronize()
                self.cuda_graph = False
                self._graphs.clear()
                return fn(*args)
            ent = self._graphs[key] = (graph, out, _lib.launch_count() - n0)
        ent[0].replay()
        self.replayed_launches += ent[2]
        return ent[1]

    @torch.no_grad()
    def __call__(self, person_gmm, cloth, person_tom):
        """person_gmm [F,22,H,W] (agnostic+cocopose), cloth [F,3,H,W], person_tom [F,7,H,W] (agnostic+densepose);
        all f32 CUDA.  Returns (p_tryon [F,3,H,W], tryon_mask [F,1,H,W], warped_cloth [F,3,H,W])."""
        return self._graph_call("call", self._stages, (person_gmm, cloth, person_tom))

    @torch.no_grad()
    def run_host(self, person_gmm_h, cloth_h, person_tom_h):
        """Host (pinned) tensors in, host (pinned) p_tryon out; the H2D / D2H copies are part of the call.

        Asynchronous and double-buffered: the copies run on two side streams so the next call's H2D overlaps this
        call's kernels.  Returns (out_host, done_event); `out_host` is valid once `done_event` has completed (or after
        a device synchronize) and is reused by the call after next.
        """
        return self._run_staged("host", (person_gmm_h, cloth_h, person_tom_h), self._stages)

    # keys of the reference's dataset batch the two stages read (datasets/tryon_dataset.py:47-61; WarpModel person
    # inputs agnostic+cocopose, UnetMaskModel person inputs agnostic+densepose, cloth for both)
    BATCH_KEYS = ("agnostic", "cocopose", "densepose", "cloth")

    @torch.no_grad()
    def run_host_batch(self, batch_h):
        """The same call on a host batch dict keyed like the reference's dataset samples (`agnostic` [F,4,H,W],
        `cocopose` [F,18,H,W], `densepose` [F,3,H,W], `cloth` [F,3,H,W], pinned f32).  Every key crosses PCIe once
        (agnostic feeds both stages: 28 channels per frame instead of 22 + 3 + 7); the per-stage channel concatenation
        is done on the device like base_model.get_and_cat_inputs (util/__init__.py:64-66)."""
        return self._run_staged("batch", tuple(batch_h[k] for k in self.BATCH_KEYS), self._stages_batch)

    def _stages_batch(self, agnostic, cocopose, densepose, cloth):
        return self._stages(torch.cat([agnostic, cocopose], 1), cloth, torch.cat([agnostic, densepose], 1))

    RAW_KEYS = ("parse", "cloth", "densepose", "image")

    @torch.no_grad()
    def run_host_raw(self, raw_h, prep):
        """Decoded 8-bit frames in (pinned uint8 host tensors, channel-last: `image`, `cloth`, `densepose` [F,H,W,3],
        `parse` [F,H,W]), host p_tryon out.  The reference's Dataset.__getitem__ tensor prep (ops.FramePrep, bit-exact)
        runs on the device, so a frame crosses PCIe as 10 bytes/pixel instead of 112."""
        if getattr(self, "_raw_prep", None) is not prep:
            self._raw_prep = prep

            def stages(parse, cloth, densepose, image):
                b = prep(parse, cloth, densepose, image)
                return self._stages(torch.cat([b["agnostic"], b["cocopose"]], 1), b["cloth"],
                                    torch.cat([b["agnostic"], b["densepose"]], 1))

            self._raw_stages = stages
        self._out_hw = raw_h["parse"].shape[1:3]
        try:
            return self._run_staged("raw", tuple(raw_h[k] for k in self.RAW_KEYS), self._raw_stages)
        finally:
            self._out_hw = None

    def _run_staged(self, tag, host_tensors, fn):
        dev = next(self.tom_model.parameters()).device
        cur = torch.cuda.current_stream(dev)
        shapes = tuple((tuple(t.shape), t.dtype) for t in host_tensors)
        st = self._host_state
        if st is None or st["shapes"] != shapes:
            frames = host_tensors[0].shape[0]
            hw = tuple(self._out_hw) if getattr(self, "_out_hw", None) else tuple(host_tensors[0].shape[2:])
            st = self._host_state = dict(
                shapes=shapes, call=0, s_in=torch.cuda.Stream(dev), s_out=torch.cuda.Stream(dev),
                dev_in=[tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host_tensors) for _ in range(2)],
                in_free=[torch.cuda.Event() for _ in range(2)], in_ready=[torch.cuda.Event() for _ in range(2)],
                out_done=[torch.cuda.Event() for _ in range(2)],
                host_out=[torch.empty((frames, 3) + hw, dtype=torch.float32).pin_memory() for _ in range(2)])
            for e in st["in_free"] + st["out_done"]:
                e.record(cur)
        slot = st["call"] % 2
        st["call"] += 1
        dev_in = st["dev_in"][slot]
        with torch.cuda.stream(st["s_in"]):
            st["s_in"].wait_event(st["in_free"][slot])  # kernels of the call before last have consumed this slot
            for d, h in zip(dev_in, host_tensors):
                d.copy_(h, non_blocking=True)
            st["in_ready"][slot].record(st["s_in"])
        cur.wait_event(st["in_ready"][slot])
        if self.cuda_graph:
            cur.wait_event(st["out_done"][slot])  # this slot's graph owns its output buffer: the last D2H of it must be done
        p_tryons, _, _ = self._graph_call(tag, fn, dev_in)
        st["in_free"][slot].record(cur)
        computed = torch.cuda.Event()
        computed.record(cur)
        out_h = st["host_out"][slot]
        if not self.cuda_graph:
            p_tryons.record_stream(st["s_out"])
        with torch.cuda.stream(st["s_out"]):
            st["s_out"].wait_event(computed)
            out_h.copy_(p_tryons, non_blocking=True)
            st["out_done"][slot].record(st["s_out"])
        return out_h, st["out_done"][slot]

    def host_sync(self):
        """Wait for every outstanding run_host call (copies included)."""
        st = self._host_state
        if st is not None:
            st["s_in"].synchronize()
            st["s_out"].synchronize()
        torch.cuda.current_stream().synchronize()
