"""FlowNet — FlowNet2 wrapper producing flow + confidence (reference: models/flownet.py:11-63).

The reference constructor loads `models/flownet2_pytorch/FlowNet2_checkpoint.pth.tar` and calls `.cuda()`; no
checkpoint is available offline, so `checkpoint=None` builds the same module tree with its default initialisation
(load one with `load_state_dict` — the keys are the reference's)."""
import os

import torch
from torch import nn

from .. import ops
from ..networks.flownet2.native_ops import Resample2d
from ..networks.flownet2.nets import FlowNet2


class FlowNet(nn.Module):
    MAX_GRAPHS = 8  # captured forwards kept (one per input-buffer pair and lane)

    def __init__(self, checkpoint=None):
        super().__init__()
        self.flowNet = FlowNet2()
        if checkpoint is not None:
            state = torch.load(checkpoint, map_location="cpu")
            self.flowNet.load_state_dict(state["state_dict"] if "state_dict" in state else state)
        self.flowNet.eval()
        self.resample = Resample2d()
        self.downsample = torch.nn.AvgPool2d(3, stride=2, padding=[1, 1], count_include_pad=False)
        # cuda_graph = True: the ~250 launches of a forward (most of them tens of microseconds at the try-on batch sizes:
        # launch-bound) are captured once per pair of input buffers and replayed; the returned tensors are then the graph's
        # static outputs, overwritten by the next call on the same input buffers.
        self.cuda_graph = False
        self._graphs = {}
        # Compute lanes (cuda_graph mode): forward(..., lane=i) replays on stream i % lanes, so two independent pair batches
        # are in flight and the many small launches of one (pyramid levels of 4x3 ... 16x12 pixels) fill the SMs the other
        # leaves idle.  The results are then ordered on the lane: join_lanes() before reading them on the caller's stream.
        self.lanes = int(os.environ.get("SHINEON_FLOW_LANES", "2"))
        self._lane_streams = {}

    def forward(self, input_A, input_B, lane=None):
        with torch.no_grad():
            size = input_A.size()
            assert len(size) == 4 or len(size) == 5
            if len(size) == 5:
                b, n, c, h, w = size
                flow, conf = self.compute_flow_and_conf(input_A.reshape(-1, c, h, w), input_B.reshape(-1, c, h, w), lane)
                return flow.view(b, n, 2, h, w), conf.view(b, n, 1, h, w)
            return self.compute_flow_and_conf(input_A, input_B, lane)

    def lane_stream(self, lane, device):
        """The stream forward(..., lane=lane) issues its work on (callers queue the download of a lane's result there)."""
        key = (str(torch.device(device)), 1 + lane % max(1, self.lanes))
        ls = self._lane_streams.get(key)
        if ls is None:
            ls = self._lane_streams[key] = torch.cuda.Stream(device)
        return ls

    def join_lanes(self):
        """Makes the caller's stream wait for every forward issued on a compute lane."""
        cur = torch.cuda.current_stream()
        for st in self._lane_streams.values():
            cur.wait_stream(st)

    def compute_flow_and_conf(self, im1, im2, lane=None):
        if not self.cuda_graph or ops.PROFILE is not None:
            return self._compute_flow_and_conf(im1, im2)
        im1, im2 = im1.contiguous(), im2.contiguous()
        lane_id = 0 if (lane is None or self.lanes <= 1) else 1 + lane % self.lanes
        key = (im1.data_ptr(), im2.data_ptr(), tuple(im1.shape), lane_id)
        if lane_id:
            cur = torch.cuda.current_stream(im1.device)
            ls = self.lane_stream(lane, im1.device)
            ls.wait_stream(cur)
            with torch.cuda.stream(ls):
                return self._replay(key, im1, im2, lane_id)
        return self._replay(key, im1, im2, lane_id)

    def _replay(self, key, im1, im2, lane_id):
        from ..networks.flownet2 import nets

        ent = self._graphs.get(key)
        if ent is None:
            while len(self._graphs) >= self.MAX_GRAPHS:
                self._graphs.pop(next(iter(self._graphs)))  # oldest capture first
            nets.CONCAT_LANE[0] = lane_id  # concurrent forwards keep separate concat buffers
            try:
                for _ in range(2):  # weight packing / allocator warm-up outside the capture
                    self._compute_flow_and_conf(im1, im2)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):  # other threads (NCCL watchdog) may poll events
                    out = self._compute_flow_and_conf(im1, im2)
            finally:
                nets.CONCAT_LANE[0] = 0
            # the graph holds raw pointers into the sub-networks' cached concat buffers: keep those tensors alive with it
            # (the caches drop their entries when the batch size / resolution changes)
            keep = [c for m in self.flowNet.modules() for c in m.__dict__.get("_cat_cache", {}).values()]
            ent = self._graphs[key] = (graph, out, keep)
        ent[0].replay()
        return ent[1]

    def _compute_flow_and_conf(self, im1, im2):
        assert im1.size()[1] == 3
        assert im1.size() == im2.size()
        old_h, old_w = im1.size()[2], im1.size()[3]
        new_h, new_w = old_h // 64 * 64, old_w // 64 * 64
        resized = old_h != new_h  # as written in the reference: only the height decides (flownet.py:47)
        if resized:
            if new_h == 0 or new_w == 0:
                raise ValueError(f"FlowNet needs inputs of at least 64x64, got {old_h}x{old_w}")
            im1 = ops.bilinear_resize(im1.contiguous(), (new_h, new_w))
            im2 = ops.bilinear_resize(im2.contiguous(), (new_h, new_w))
        elif old_w != new_w:
            raise ValueError(
                f"FlowNet: width {old_w} is not a multiple of 64 while the height is; the reference does not resize in "
                "that case either (flownet.py:47 tests old_h only) and its FlowNet2 then fails on mismatched pyramid levels")
        data1 = torch.stack([im1, im2], dim=2).contiguous()  # [B,3,2,H,W]
        flow1 = self.flowNet(data1)
        conf = ops.flow_confidence(im1.contiguous(), im2.contiguous(), flow1, 0.02)
        if resized:
            # flownet.py:56-58: `upsample(flow1) * old_h / new_h` (both flow components scaled by the height ratio)
            flow1 = ops.bilinear_resize(flow1.contiguous(), (old_h, old_w), mul=1.0)
            flow1 = flow1 * old_h / new_h
            conf = ops.bilinear_resize(conf.contiguous(), (old_h, old_w))
        return flow1.detach(), conf.detach()
