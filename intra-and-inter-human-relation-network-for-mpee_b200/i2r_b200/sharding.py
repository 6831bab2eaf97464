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


# ---------------------------------------------------------------------------------------------------------------
# Partitioning (2) of SURVEY.md 8e: crops sharded across ranks, ONE all-gather of the pooled token maps.
#
# Rank r owns the contiguous crop slice [r*S_max, min((r+1)*S_max, S)), S_max = ceil(S / world); slices are padded to
# S_max crops with zeros, so the gathered buffer [world * S_max, tokens, channels] is in global crop order with the
# padding at the very end.  After the gather every rank runs the inter-human encoder on a WINDOW of crops around its
# slice -- wide enough to hold every image that has a crop in the slice whole -- and finishes (upsample, residual, head)
# its own crops only.  Images straddling a slice boundary are encoded on both sides (the inter-human stage is 0.6-2 % of
# the FLOPs); nothing has to be sent back, so the all-gather is the only collective of the forward.
class CropShardLayout:
    """Capture-time constants of one rank: slice, window and the offset of the slice inside the window.  They depend on
    (S, world, rank, persons_bound) only -- never on the persons-per-image list, which stays device data."""

    def __init__(self, total_crops, world_size, rank, persons_bound):
        s, w = int(total_crops), int(world_size)
        self.total, self.world, self.rank = s, w, int(rank)
        self.persons_bound = min(int(persons_bound), s)
        self.s_max = (s + w - 1) // w
        self.s_pad = self.s_max * w
        self.c0 = min(self.rank * self.s_max, s)
        self.c1 = min(self.c0 + self.s_max, s)
        margin = self.persons_bound - 1                     # an image reaches at most this far beyond the slice
        self.wn = min(self.s_pad, self.s_max + 2 * margin)
        self.w0 = max(0, min(self.rank * self.s_max - margin, self.s_pad - self.wn))
        self.local_offset = self.rank * self.s_max - self.w0      # first local crop inside the window

    def window_lengths(self, length):
        """Sequence structure of the window: the images clipped to it (an image with a crop in this rank's slice is never
        clipped, given persons <= persons_bound), padding crops past S as singleton sequences.  Sums to `wn`."""
        if max(length) > self.persons_bound:
            raise ValueError("an image has %d persons; this sharded forward was built for at most %d" % (
                max(length), self.persons_bound))
        if sum(length) != self.total:
            raise ValueError("sum(length)=%d, expected %d crops" % (sum(length), self.total))
        out, a = [], 0
        lo, hi = self.w0, self.w0 + self.wn
        for n in length:
            b = a + n
            k = min(b, hi) - max(a, lo)
            if k > 0:
                out.append(k)
            a = b
        out += [1] * max(0, hi - max(lo, self.total))
        assert sum(out) == self.wn
        return out

    def owned_images_whole(self, length):
        """True when every image with a crop in [c0, c1) lies inside the window (checked by the tests)."""
        a = 0
        for n in length:
            b = a + n
            if b > self.c0 and a < self.c1 and not (a >= self.w0 and b <= self.w0 + self.wn):
                return False
            a = b
        return True
