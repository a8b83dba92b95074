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
