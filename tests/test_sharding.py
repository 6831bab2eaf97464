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


# ---------------------------------------------------------------------------------------------------------------
# Crop-sharded forward with one all-gather (i2r_b200/sharded.py), host logic on CPU: two gloo ranks run the product's
# launch sequence through the emulator (tests/emulator.py) on their crop slices, exchange the pooled token maps with
# ONE all_gather_into_tensor and must reproduce the single-process forward.
def test_crop_shard_layout_windows_hold_owned_images_whole():
    import random
    from i2r_b200.sharding import CropShardLayout
    rng = random.Random(1)
    for _ in range(500):
        world = rng.choice([1, 2, 3, 4, 8])
        length = [rng.randint(1, 6) for _ in range(rng.randint(1, 12))]
        s = sum(length)
        pb = rng.choice([max(length), max(length) + 1, s])
        covered = []
        for r in range(world):
            lay = CropShardLayout(s, world, r, pb)
            wl = lay.window_lengths(length)
            assert sum(wl) == lay.wn and max(wl) <= lay.persons_bound and lay.owned_images_whole(length)
            assert 0 <= lay.local_offset and lay.local_offset + lay.s_max <= lay.wn
            covered += list(range(lay.c0, lay.c1))
        assert covered == list(range(s))
    lay = CropShardLayout(64, 8, 3, 8)           # C4: 8 images x 8 persons over 8 ranks -> one image per rank
    assert (lay.c0, lay.c1, lay.wn, lay.w0, lay.local_offset) == (24, 32, 22, 17, 7)
    assert lay.window_lengths([8] * 8) == [7, 8, 7]


def _sharded_worker(rank, world, port, length, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(paths.REPO, "tests"))
    from emulator import EmuRunner
    from helpers import build_model, inputs_for
    from i2r_b200.sharded import ShardedForward
    torch.set_num_threads(2)
    cfg, model, _ = build_model()
    model._runner_factory = lambda device, impl: EmuRunner()
    model.prepare("cpu")
    x, pm = inputs_for(length)
    sf = ShardedForward(model, persons_bound=max(length), gather_output=True)
    sf.use_cuda_graph = False
    with torch.no_grad():
        full = sf._run(x, pm, length, torch.device("cpu"))
        sf.gather_output = False
        local = sf._run(x, pm, length, torch.device("cpu"))
    c0, c1 = sf.crop_range(sum(length))
    ref = None
    if rank == 0:
        with torch.no_grad():
            ref = model._eager(x, pm, length)
    out_q.put((rank, full.numpy().copy(), local.numpy().copy(), (c0, c1), None if ref is None else ref.numpy().copy(),
               sf.bytes_gathered))
    dist.destroy_process_group()


def test_two_rank_gloo_crop_sharded_forward_matches_single_process():
    length = [2, 1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, length, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = res[0][4]
    assert ref is not None and tuple(ref.shape) == (3, 17, 64, 48)
    for rank, full, local, (c0, c1), _, nbytes in res:
        assert tuple(full.shape) == tuple(ref.shape)
        assert float(abs(full - ref).max()) <= 1e-5             # same launches on the same crops
        assert (local == full[c0:c1]).all()
        assert nbytes == 2 * 2 * 2 * 192 * 96 * 2                # world x s_max x (tokens, pos) x 192 tokens x 96 ch fp16
    assert [r[3] for r in res] == [(0, 2), (2, 3)]
