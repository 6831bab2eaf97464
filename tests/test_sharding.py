"""CPU tests of the multi-GPU host logic with a real 2-process gloo group: image sharding covers every crop
exactly once, gathers back in order, and the timing reduction takes the max over ranks."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import paths  # noqa: F401
from i2r_b200.sharding import crop_slices, max_over_ranks, shard_images, take_shard


def test_shard_images_balances_and_partitions():
    length = [4, 1, 3, 2, 2, 1, 5, 1]
    for world in (1, 2, 3, 4, 8):
        owned = shard_images(length, world)
        flat = sorted(i for o in owned for i in o)
        assert flat == list(range(len(length)))
        loads = [sum(length[i] for i in o) for o in owned]
        assert max(loads) - min(loads) <= max(length)
    assert crop_slices([2, 1, 3]) == [(0, 2), (2, 3), (3, 6)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, length, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    s = sum(length)
    x = torch.randn(s, 3, 8, 6, generator=g)
    pm = torch.rand(s, 1, 8, 6, generator=g)
    owned = shard_images(length, world)[rank]
    xs, pms, ls, idx = take_shard(x, pm, length, owned)
    assert xs.shape[0] == sum(ls) == len(idx)
    # stand-in for the per-rank forward: a per-crop function (the real path is per-image independent)
    local = xs.mean(dim=(1, 2, 3)) + pms.mean(dim=(1, 2, 3))
    gathered = [None] * world
    dist.all_gather_object(gathered, (idx, local))
    full = torch.empty(s)
    seen = []
    for ids, vals in gathered:
        full[torch.tensor(ids)] = vals
        seen += ids
    ref = x.mean(dim=(1, 2, 3)) + pm.mean(dim=(1, 2, 3))
    mx = max_over_ranks([float(rank + 1), 10.0 - rank])
    out_q.put((rank, sorted(seen) == list(range(s)), bool(torch.allclose(full, ref)), mx))
    dist.destroy_process_group()


def test_two_rank_gloo_shard_gather_and_max_timing():
    length = [3, 1, 2, 2, 1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, length, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, covered, equal, mx in res:
        assert covered and equal
        assert mx == [2.0, 10.0]
