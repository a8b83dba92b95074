"""World-size-2 gloo test of the N>1 host logic (clip sharding, barrier, max-over-ranks timing reduction)."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from shineon_virtual_tryon_b200 import distributed as d

    r, w = d.init_process_group("gloo")
    assert (r, w) == (rank, world)
    mine = d.shard_clips(n_clips, r, w)
    d.barrier()
    slowest = d.max_over_ranks(10.0 + rank)          # rank 1 is "slower"
    total = d.sum_over_ranks(len(mine))
    out.put((rank, mine, slowest, total))
    torch.distributed.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    world, n_clips = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = [r[1] for r in res]
    assert sorted(shards[0] + shards[1]) == list(range(n_clips))      # every clip exactly once
    assert abs(len(shards[0]) - len(shards[1])) <= 1                   # balanced
    assert all(abs(r[2] - 11.0) < 1e-9 for r in res)                   # max over ranks
    assert all(abs(r[3] - n_clips) < 1e-9 for r in res)               # whole-job count


def test_single_process_is_a_noop():
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from shineon_virtual_tryon_b200 import distributed as d

    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        os.environ.pop(k, None)
    assert d.init_process_group("gloo") == (0, 1)
    assert d.shard_clips(5, 0, 1) == [0, 1, 2, 3, 4]
    assert d.max_over_ranks(3.5) == 3.5


def _grad_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from shineon_virtual_tryon_b200 import distributed as d

    d.init_process_group("gloo")
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2, 2))]
    red = d.FlatGradAllReducer(params, bucket_bytes=32)  # tiny buckets -> several async all-reduces
    for i in range(len(params)):
        red.grad_view(i).fill_(float(rank + 1) * (i + 1))
    red.start()
    scale = red.finish()
    out.put((rank, [(red.grad_view(i) * scale).flatten()[0].item() for i in range(len(params))], len(red.buckets)))
    torch.distributed.destroy_process_group()


def test_flat_grad_allreduce_is_mean_over_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, means, nb in res:
        assert nb > 1
        assert means == [1.5 * (i + 1) for i in range(3)]  # mean of (rank+1)*(i+1) over ranks {0,1}


def _overlap_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from shineon_virtual_tryon_b200 import distributed as d

    d.init_process_group("gloo")
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 11, 3, 8, 6)]
    red = d.FlatGradAllReducer(params, bucket_bytes=4 * 8)  # 8-element buckets: 33 elements -> 5 buckets
    red.begin_overlap()
    launched = []
    # the backward pass reports parameters in buffer order (training.backward_order); buckets go out as they fill
    for i, p in enumerate(params):
        red.grad_view(i).fill_(float((rank + 1) * (i + 1)))
        red.mark_ready([p])
        launched.append(red._next_bucket)
    scale = red.end_overlap()
    vals = [(red.grad_view(i) * scale).unique().tolist() for i in range(len(params))]
    out.put((rank, launched, vals, len(red.buckets)))
    torch.distributed.destroy_process_group()


def test_overlapped_bucket_launch_order_and_result():
    """Row U7: buckets are all-reduced as soon as the last gradient they contain is final, and the result is the mean."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, launched, vals, nb in res:
        assert nb == 5
        # parameter ends at 5, 16, 19, 27, 33 -> complete buckets after each: 0, 2, 2, 3, 5
        assert launched == [0, 2, 2, 3, 5]
        assert vals == [[1.5 * (i + 1)] for i in range(5)]


def test_backward_order_covers_every_unet_parameter_once():
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from shineon_virtual_tryon_b200.models.unet_mask_model import UnetMaskModel
    from shineon_virtual_tryon_b200.training import backward_order
    from tests.util import make_hparams

    m = UnetMaskModel(make_hparams(is_train=True))
    order = backward_order(m.unet)
    assert len({id(p) for p in order}) == len(order) == len(list(m.unet.parameters()))
    # first finalised: the outermost up-conv (the last layer of the forward); last: the outermost down-conv
    assert order[0] is m.unet.model._parts["upconv"].weight and order[-2] is m.unet.model._parts["downconv"].weight
