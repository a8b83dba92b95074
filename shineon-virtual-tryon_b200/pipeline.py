"""In-memory two-stage try-on: GMM -> TPS warp -> TOM, frames batched along N.

The reference runs the two stages as separate `test.py` invocations and passes the warped cloth through 8-bit
PNGs on disk (docs/2_inference.md:13-39, models/warp_model.py:146-148, datasets/vvt_dataset.py:133-150).
This class chains the same two forwards (WarpModel.forward + grid_sample(border), UnetMaskModel.forward)
on the device.  Entry points, from the reference's tensors to its file formats:

  pipe(person_gmm, cloth, person_tom)      f32 device tensors in, f32 device tensors out (the nn.Module surface)
  pipe.run_raw(parse, cloth, densepose, image, prep)
                                           decoded 8-bit frames on the device in (what Dataset.__getitem__ starts from),
                                           8-bit try-on frames out (what visualization.save_images hands to the PNG encoder)
  pipe.run_host / run_host_batch           pinned f32 host tensors in, pinned f32 host image out
  pipe.run_host_raw(raw_h, prep)           pinned uint8 host frames in, pinned uint8 host frames out: 10 B/pixel up,
                                           3 B/pixel down (the f32 forms move 112 and 12)
"""
import os
from collections import OrderedDict

import torch


class TryOnPipeline:
    MAX_GRAPHS = 8

    def __init__(self, warp_model, tom_model, cuda_graph=False):
        """cuda_graph=True: the ~180 launches of a step are captured once per distinct set of input buffers and
        replayed (the small launches at the bottom of the U-Net are otherwise issued slower than the GPU runs them).
        Replayed calls return the graph's static output tensors: they are overwritten by the next call with the same
        input buffers.  A graph is keyed on its input buffers, the stage function and the models' weights / precision,
        so loading a checkpoint or changing a parameter re-captures instead of replaying stale packed weights."""
        self.warp_model = warp_model
        self.tom_model = tom_model
        self._host_state = {}
        self.cuda_graph = bool(cuda_graph)
        self._graphs = OrderedDict()
        self._tensors = None
        self.replayed_launches = 0  # kernel launches executed through graph replays (not seen by shineon_launch_count)
        # Compute lanes: consecutive steps are independent (frames / clips never talk to each other), so the host-buffer
        # entry points and run_raw(lane=...) replay step i on stream i % lanes.  Two steps in flight let the launch-latency
        # bound parts of one step (the 4x3 ... 16x12 levels, attention, the regression head: kernels of 8-40 CTAs) run next
        # to the chip-filling layers of the other.  1 (default) = everything on the caller's stream: measured on B200 at 160
        # frames per step, two lanes gave +1 % device-resident and -4 % end to end (profiles/r02_concurrency.md).
        self.lanes = int(os.environ.get("SHINEON_LANES", "1"))
        self._lane_streams = {}

    def set_precision(self, precision):
        self.warp_model.set_precision(precision)
        self.tom_model.set_precision(precision)
        self._graphs.clear()  # captured graphs hold the previous mode's kernels

    # ------------------------------------------------------------------ stage functions (device tensors -> device tensors)
    def _stages(self, person_gmm, cloth, person_tom):
        warped_cloth, _, _, _ = self.warp_model.warp(person_gmm, cloth, cloth)
        _, tryon_masks, p_tryons, _ = self.tom_model(person_tom, warped_cloth)
        return p_tryons, tryon_masks, warped_cloth

    def _stages_batch(self, agnostic, cocopose, densepose, cloth):
        return self._stages(torch.cat([agnostic, cocopose], 1), cloth, torch.cat([agnostic, densepose], 1))

    def _raw_fn(self, prep, u8_out):
        """Stage function over decoded 8-bit frames for one FramePrep instance (cached: graphs are keyed on it)."""
        cache = self.__dict__.setdefault("_raw_fns", {})
        key = (id(prep), bool(u8_out))
        if key not in cache:
            def stages_fused(parse, cloth, densepose, image):
                # frame prep writes the three stem operands directly; the sampler reads the 8-bit cloth and fills the
                # U-Net operand's cloth channels: no f32 dataset tensors, no torch.cat, no layout passes
                from . import ops

                b = ops.frame_prep_planes(prep, parse, cloth, densepose, image, prec=ops.resolve_precision(self.warp_model.precision))
                theta = self.warp_model.regress_theta(b["gmm_person"], b["cloth_i2c"])
                warped = self.warp_model.gridGen.warp_u8(theta, cloth, b["unet_in"])
                u8 = self.tom_model.forward_u8_planes(b["unet_in"], warped)
                return (u8.view(u8.shape[0], u8.shape[2], u8.shape[3], 3),)

            def stages(parse, cloth, densepose, image):
                if u8_out and self.FUSED_PREP and self._fused_prep_ok(prep):
                    return stages_fused(parse, cloth, densepose, image)
                b = prep(parse, cloth, densepose, image)
                person_gmm = torch.cat([b["agnostic"], b["cocopose"]], 1)
                person_tom = torch.cat([b["agnostic"], b["densepose"]], 1)
                if not u8_out:
                    return self._stages(person_gmm, b["cloth"], person_tom)
                warped_cloth, _, _, _ = self.warp_model.warp(person_gmm, b["cloth"], b["cloth"])
                u8 = self.tom_model.forward_u8(person_tom, warped_cloth)  # [F,1,H,W,3]
                return (u8.view(u8.shape[0], u8.shape[2], u8.shape[3], 3),)

            cache[key] = (stages, prep)  # keeps `prep` alive so its id stays unique
        return cache[key][0]

    # True: run_raw / run_host_raw(u8_out=True) use the fused frame-prep -> stem-operand path (ops.frame_prep_planes);
    # False keeps prep -> f32 tensors -> torch.cat -> layout kernels (same results bit for bit; A/B and tests)
    FUSED_PREP = True

    def _fused_prep_ok(self, prep):
        """The fused path covers the reference recipe's stems: WarpModel on agnostic+cocopose / 3-channel cloth, U-Net on
        agnostic+densepose+cloth (10 channels), one frame per sample, same precision in both models."""
        from . import ops
        from .networks.cpvton.warp import FeatureExtraction  # noqa: F401

        w, t = self.warp_model, self.tom_model
        try:
            ok = (list(w.hparams.person_inputs) == ["agnostic", "cocopose"] and list(t.hparams.person_inputs) == ["agnostic", "densepose"]
                  and list(w.hparams.cloth_inputs) == ["cloth"] and list(t.hparams.cloth_inputs) == ["cloth"]
                  and w.extractionA.model[0].in_channels == 4 + prep.J and w.extractionB.model[0].in_channels == 3
                  and t.unet.model._parts["downconv"].in_channels == 10 and t.hparams.n_frames_total == 1
                  and ops.resolve_precision(w.precision) == ops.resolve_precision(t.unet.precision)
                  and w.gridGen.grid_size in (3, 5))
        except AttributeError:
            return False
        return ok

    # ------------------------------------------------------------------ CUDA-graph cache
    def _weights_signature(self):
        if self._tensors is None:
            self._tensors = [t for m in (self.warp_model, self.tom_model) for t in list(m.parameters()) + list(m.buffers())]
        from .networks._engine_util import WEIGHTS_EPOCH

        return (WEIGHTS_EPOCH[0],) + tuple((t.data_ptr(), t._version) for t in self._tensors)

    def _graph_call(self, tag, fn, args):
        """fn(*args) through a CUDA graph keyed by (tag, fn, argument buffers, weights).  Eager while ops.PROFILE is
        collecting."""
        from . import _lib, ops

        if not self.cuda_graph or ops.PROFILE is not None:
            return fn(*args)
        # bound methods are created anew on every attribute access: key them on the underlying function, closures on id
        fn_id = id(getattr(fn, "__func__", fn))
        key = (tag, fn_id, self._weights_signature()) + tuple((a.data_ptr(), tuple(a.shape), a.dtype) for a in args)
        ent = self._graphs.get(key)
        if ent is None:
            while len(self._graphs) >= self.MAX_GRAPHS:
                self._graphs.popitem(last=False)  # least recently used
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
        else:
            self._graphs.move_to_end(key)
        ent[0].replay()
        self.replayed_launches += ent[2]
        return ent[1]

    # ------------------------------------------------------------------ device entry points
    @torch.no_grad()
    def __call__(self, person_gmm, cloth, person_tom):
        """person_gmm [F,22,H,W] (agnostic+cocopose), cloth [F,3,H,W], person_tom [F,7,H,W] (agnostic+densepose);
        all f32 CUDA.  Returns (p_tryon [F,3,H,W], tryon_mask [F,1,H,W], warped_cloth [F,3,H,W])."""
        return self._graph_call("call", self._stages, (person_gmm, cloth, person_tom))

    RAW_KEYS = ("parse", "cloth", "densepose", "image")

    def _lane(self, i, dev):
        key = (str(dev), i % max(1, self.lanes))
        st = self._lane_streams.get(key)
        if st is None:
            st = self._lane_streams[key] = torch.cuda.Stream(dev)
        return st

    def join_lanes(self):
        """Makes the caller's stream wait for every step issued on a compute lane (run_raw(lane=...))."""
        cur = torch.cuda.current_stream()
        for st in self._lane_streams.values():
            cur.wait_stream(st)

    @torch.no_grad()
    def run_raw(self, parse, cloth, densepose, image, prep, u8_out=True, lane=None):
        """Decoded 8-bit frames on the device (`image`, `cloth`, `densepose` [F,H,W,3], `parse` [F,H,W], uint8) ->
        try-on frames uint8 [F,H,W,3] exactly as visualization.save_images would encode them (u8_out=False: the f32
        triple of __call__).  The reference's Dataset.__getitem__ tensor prep runs on the device (ops.FramePrep).

        lane=i (CUDA-graph mode): the step is issued on compute lane i % self.lanes after everything queued on the caller's
        stream so far; steps on different lanes overlap.  The result is ordered on the lane, not on the caller's stream:
        call join_lanes() (or synchronise) before reading it, and give concurrent steps distinct input buffers (a graph
        owns its output buffer)."""
        fn, args = self._raw_fn(prep, u8_out), (parse, cloth, densepose, image)
        if lane is None or self.lanes <= 1 or not self.cuda_graph:
            out = self._graph_call("raw", fn, args)
        else:
            cur = torch.cuda.current_stream(parse.device)
            ls = self._lane(lane, parse.device)
            ls.wait_stream(cur)
            with torch.cuda.stream(ls):
                out = self._graph_call("raw", fn, args)
        return out[0] if u8_out else out

    # ------------------------------------------------------------------ host entry points
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

    @torch.no_grad()
    def run_host_raw(self, raw_h, prep, u8_out=True):
        """Decoded 8-bit frames in (pinned uint8 host tensors, channel-last: `image`, `cloth`, `densepose` [F,H,W,3],
        `parse` [F,H,W]), try-on frames out: pinned uint8 [F,H,W,3] (the bytes the reference's PNG writer receives), or
        the f32 p_tryon [F,3,H,W] with u8_out=False.  A frame crosses PCIe as 10 bytes/pixel up and 3 down instead of
        112 and 12."""
        return self._run_staged("raw_u8" if u8_out else "raw", tuple(raw_h[k] for k in self.RAW_KEYS), self._raw_fn(prep, u8_out))

    def _run_staged(self, tag, host_tensors, fn):
        dev = next(self.tom_model.parameters()).device
        cur = torch.cuda.current_stream(dev)
        shapes = tuple((tuple(t.shape), t.dtype) for t in host_tensors)
        st = self._host_state.get(tag)
        if st is None or st["shapes"] != shapes:
            st = self._host_state[tag] = dict(
                shapes=shapes, call=0, s_in=torch.cuda.Stream(dev), s_out=torch.cuda.Stream(dev),
                dev_in=[tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host_tensors) for _ in range(2)],
                in_free=[torch.cuda.Event() for _ in range(2)], in_ready=[torch.cuda.Event() for _ in range(2)],
                out_done=[torch.cuda.Event() for _ in range(2)], host_out=[None, None])
            for e in st["in_free"] + st["out_done"]:
                e.record(cur)
        self._last_state = st
        slot = st["call"] % 2
        st["call"] += 1
        dev_in = st["dev_in"][slot]
        with torch.cuda.stream(st["s_in"]):
            st["s_in"].wait_event(st["in_free"][slot])  # kernels of the call before last have consumed this slot
            for d, h in zip(dev_in, host_tensors):
                d.copy_(h, non_blocking=True)
            st["in_ready"][slot].record(st["s_in"])
        # compute stream of this call: the caller's, or (CUDA-graph mode, lanes > 1) the slot's lane so that two
        # consecutive calls' kernels overlap; either way the result is ordered by `done_event`
        comp = cur
        if self.cuda_graph and self.lanes > 1:
            comp = self._lane(slot, dev)
            comp.wait_stream(cur)
        comp.wait_event(st["in_ready"][slot])
        if self.cuda_graph:
            comp.wait_event(st["out_done"][slot])  # this slot's graph owns its output buffer: the last D2H of it must be done
        with torch.cuda.stream(comp):
            result = self._graph_call(tag, fn, dev_in)[0]
        st["in_free"][slot].record(comp)
        computed = torch.cuda.Event()
        computed.record(comp)
        if st["host_out"][slot] is None:
            st["host_out"][slot] = torch.empty(result.shape, dtype=result.dtype).pin_memory()
        out_h = st["host_out"][slot]
        if not self.cuda_graph:
            result.record_stream(st["s_out"])
        with torch.cuda.stream(st["s_out"]):
            st["s_out"].wait_event(computed)
            out_h.copy_(result, non_blocking=True)
            st["out_done"][slot].record(st["s_out"])
        return out_h, st["out_done"][slot]

    def host_streams(self):
        """The copy streams of every host entry point used so far (a caller timing on its own stream waits on these)."""
        return [s for st in self._host_state.values() for s in (st["s_in"], st["s_out"])] + list(self._lane_streams.values())

    def host_sync(self):
        """Wait for every outstanding run_host* call (copies included)."""
        for s in self.host_streams():
            s.synchronize()
        for s in self._lane_streams.values():
            s.synchronize()
        torch.cuda.current_stream().synchronize()
