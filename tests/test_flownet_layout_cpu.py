"""Host-side layout logic of the FlowNet2 engine that needs no GPU: concat-buffer segment packing, the channel maps the
consumers' packed weights follow, and the channel windows producers write into (networks/flownet2/nets.py, ops.Planes)."""
import pytest
import torch

from shineon_virtual_tryon_b200 import ops
from shineon_virtual_tryon_b200.networks.flownet2 import nets


def test_chan_map_one_kblock_per_segment():
    """Default layout (FlowNetC / S / SD): every segment starts on a 64-channel K-block (FlowNetS.py:80-92 concat orders)."""
    cm = nets._Concat.chan_map([128, 64, 2])
    assert len(cm) == 256
    assert cm[:128] == list(range(128)) and cm[128:192] == list(range(128, 192))
    assert cm[192:194] == [192, 193] and set(cm[194:]) == {-1}
    cm = nets._Concat.chan_map([32, 441])  # FlowNetC in31 = [conv_redir | corr] (FlowNetC.py:90)
    assert len(cm) == 64 + 448 and cm[:32] == list(range(32)) and set(cm[32:64]) == {-1}
    assert cm[64:64 + 441] == list(range(32, 473)) and set(cm[64 + 441:]) == {-1}


def test_chan_map_packed_segments():
    """FlowNetFusion's packed layout: 64 + 16 + 2 channels in 128, 128 + 32 + 2 in 192 (FlowNetFusion.py:58-66)."""
    cm = nets._Concat.chan_map(nets._Segs((64, 16, 2), 8))
    assert len(cm) == 128 and cm[:82] == list(range(82)) and set(cm[82:]) == {-1}
    cm = nets._Concat.chan_map(nets._Segs((128, 32, 2), 8))
    assert len(cm) == 192 and cm[:162] == list(range(162)) and set(cm[162:]) == {-1}
    # segments that are not multiples of 8 get their own padding
    cm = nets._Concat.chan_map(nets._Segs((10, 3), 8))
    assert len(cm) == 64 and cm[:10] == list(range(10)) and cm[10:16] == [-1] * 6 and cm[16:19] == [10, 11, 12]


@pytest.mark.parametrize("align,offsets,width", [(64, [0, 64, 128], 192), (8, [0, 64, 80], 128)])
def test_concat_windows(align, offsets, width):
    cat = nets._Concat(2, 4, 6, (64, 16, 2), ops.resolve_precision(None), "cpu", align=align)
    assert cat.offsets == offsets and cat.buf.cpad == width and tuple(cat.buf.hi.shape) == (2, 4, 6, width)
    assert float(cat.buf.hi.abs().max()) == 0 and float(cat.buf.lo.abs().max()) == 0  # padding channels start as zeros
    for i, (off, c) in enumerate(zip(offsets, (64, 16, 2))):
        w = cat.window(i)
        assert w.coffset == off and w.C == c and w.cstride == width and w.hi is cat.buf.hi
        assert off + w.cpad <= width
    # the channel map of the consumer addresses exactly the windows' channels
    cm = nets._Concat.chan_map(cat.seg)
    assert len(cm) == width
    for i, (off, c) in enumerate(zip(offsets, (64, 16, 2))):
        base = sum((64, 16, 2)[:i])
        assert cm[off:off + c] == list(range(base, base + c))


def test_window_alignment_rules():
    p = ops.Planes(1, 2, 2, 128, device="cpu", cpad=128)
    assert p.window(64, 64).cpad == 64
    with pytest.raises(AssertionError):
        p.window(80, 2)  # a default window must start on a K-block
    w = p.window(80, 2, align=8)
    assert w.coffset == 80 and w.cpad == 8
    with pytest.raises(AssertionError):
        p.window(124, 2, align=8)  # not 8-aligned
    with pytest.raises(AssertionError):
        p.window(120, 12, align=8)  # runs past the buffer
    with pytest.raises(AssertionError):
        p.window(64, 65)  # 64-aligned windows own whole K-blocks
