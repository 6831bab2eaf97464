"""Multi-GPU partitioning of the path: whole images per rank (partitioning (1)).

Per-crop stages are independent and the inter-human encoder only couples crops of the same image
(lib/models/attention.py:131-137), so assigning whole images to ranks needs no data-path collective.
`shard_images` balances the number of crops per rank (longest-processing-time greedy); `max_over_ranks`
is the timing reduction bench.py uses (device time, max over ranks).
"""
import torch


def shard_images(length, world_size):
    """length: persons per image.  Returns, per rank, the sorted list of image indices it owns."""
    order = sorted(range(len(length)), key=lambda i: (-length[i], i))
    load = [0] * world_size
    owned = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        owned[r].append(i)
        load[r] += length[i]
    return [sorted(o) for o in owned]


def crop_slices(length):
    """Start/stop crop index of every image in the person-major batch."""
    out, s = [], 0
    for n in length:
        out.append((s, s + n))
        s += n
    return out


def take_shard(x, pos_mask, length, image_ids):
    """Sub-batch (x, pos_mask, length) of the listed images (crops stay grouped by image)."""
    sl = crop_slices(length)
    idx = [j for i in image_ids for j in range(*sl[i])]
    sel = torch.tensor(idx, dtype=torch.long)
    return x.index_select(0, sel), pos_mask.index_select(0, sel), [length[i] for i in image_ids], idx


def max_over_ranks(values, device=None):
    """Element-wise max of a list of floats over all ranks (no-op without an initialised process group)."""
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]
