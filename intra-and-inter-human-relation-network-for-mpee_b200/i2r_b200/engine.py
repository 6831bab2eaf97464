"""CUDA-graph engine around an eager kernel-launch sequence, usable under the reference's real eval loop.

The forward is a fixed sequence of ~100-600 kernel launches whose SHAPES depend only on the number of crops S and
the input size; the persons-per-image list `length` (lib/core/function.py:128 -- different on almost every batch)
enters the kernels only as DATA: the per-image token offsets `cu_seqlens` (device buffer) of the ragged attention
and two launch bounds (number of images, longest sequence).  So graphs are keyed on (S, H, W) and hold

  * a `SeqPlan`: static int32 device buffers with the token offsets, padded with empty sequences up to the plan's
    image bound, rewritten (a few hundred bytes, stream ordered) before every replay;
  * launch bounds (images <= nseq_bound, persons per image <= persons_bound) fixed at capture time.

At most TWO captures exist per (S, H, W): the first one is tight (the bounds of the `length` that triggered it --
the steady-state case of a fixed batch layout); the first `length` that does not fit captures the generic one
(nseq_bound = persons_bound = S), which fits every later list.  Entries are evicted LRU.

The sequence is captured as TWO graphs, cut at the first use of `pos_mask` (the eager function calls the
`mask_needed` hook there), and host inputs go through device staging buffers filled on a copy stream:

    copy stream :  H2D x(i+1) -> stage_x | H2D mask(i+1) -> stage_m          (overlaps the graphs of call i)
    main stream :  stage_x -> static x (D2D) | graph 1 | stage_m -> static mask (D2D) | graph 2 | clone outputs

so the upload of call i+1 overlaps the compute of call i whenever the host runs ahead (`HostPipeline` below, or
any caller that does not synchronise on every output), and the mask upload always overlaps the backbone.
"""
import collections

import torch


class SeqPlan:
    """Token offsets of a ragged batch as device data.  `cu(tokens)` is an int32 device tensor of nseq_bound + 1
    offsets (images past len(length) are empty sequences at the end), `max_seqlen(tokens)` the launch bound."""

    def __init__(self, device, total_persons, nseq_bound, persons_bound):
        self.device = torch.device(device)
        self.total = int(total_persons)
        self.nseq_bound = int(nseq_bound)
        self.persons_bound = int(persons_bound)
        self.length = None
        self._cu = {}

    @classmethod
    def exact(cls, length, device):
        plan = cls(device, sum(length), len(length), max(length))
        plan.length = [int(n) for n in length]
        return plan

    def fits(self, length):
        return len(length) <= self.nseq_bound and max(length) <= self.persons_bound and sum(length) == self.total

    def host_offsets(self, tokens):
        off = [0]
        for n in self.length:
            off.append(off[-1] + n * tokens)
        off += [off[-1]] * (self.nseq_bound + 1 - len(off))
        return torch.tensor(off, dtype=torch.int32)

    def cu(self, tokens):
        t = self._cu.get(tokens)
        if t is None:
            if self.device.type == "cuda" and torch.cuda.is_current_stream_capturing():
                raise RuntimeError("SeqPlan.cu(%d) first requested during graph capture (the warm-up run must see "
                                   "every tokens-per-person value)" % tokens)
            t = self.host_offsets(tokens).to(self.device)
            self._cu[tokens] = t
        return t

    def max_seqlen(self, tokens):
        return self.persons_bound * tokens

    def update(self, length):
        """New persons-per-image list for the same crops: rewrite the device offsets on the current stream.  The source
        is pageable host memory, so the bytes are staged by the driver at call time (no host buffer to keep alive, no
        device synchronisation)."""
        length = [int(n) for n in length]
        if not self.fits(length):
            raise ValueError("length %r does not fit this plan (crops %d, images <= %d, persons <= %d)" % (
                length, self.total, self.nseq_bound, self.persons_bound))
        if length == self.length:
            return
        self.length = length
        for tokens, buf in self._cu.items():
            buf.copy_(self.host_offsets(tokens), non_blocking=True)


class _Entry:
    __slots__ = ("plan", "sx", "sm", "stage_x", "stage_m", "segments", "out", "copy_stream", "x_ready", "mask_ready",
                 "x_stage_free", "m_stage_free")


class _CaptureHooks:
    """What the eager function sees while its launch sequence is being captured: `mask_needed()` and `between(fn)`
    end the current graph segment; the segment that follows is preceded, at every replay, by the mask hand-over or by
    `fn()` run EAGERLY on the launching stream (collectives, copies whose offsets are not capture-time constants)."""

    def __init__(self, pool):
        self.pool = pool
        self.segments = []            # [(pre, graph)], pre in (None, "mask", callable)
        self.ctx = None
        self.mask_seen = False
        self._open(None)

    def _open(self, pre):
        g = torch.cuda.CUDAGraph()
        self.ctx = torch.cuda.graph(g, pool=self.pool)
        self.ctx.__enter__()
        self.segments.append((pre, g))

    def _close(self):
        if self.ctx is not None:
            self.ctx.__exit__(None, None, None)
            self.ctx = None

    def mask_needed(self):
        if not self.mask_seen:        # cut: everything recorded so far needs the images only
            self.mask_seen = True
            self._close()
            self._open("mask")

    def between(self, fn):
        self._close()
        fn()
        self._open(fn)


class EagerHooks:
    """The same interface outside capture."""

    def mask_needed(self):
        pass

    def between(self, fn):
        fn()


class GraphedForward:
    """eager_fn(x, pos_mask, plan, hooks=None) -> tensor or dict of tensors (all launches on the current stream, every
    allocation through torch's caching allocator)."""

    def __init__(self, eager_fn, max_entries=8):
        self.eager_fn = eager_fn
        self.max_entries = max_entries
        self.entries = collections.OrderedDict()      # (shape key, bounds) -> _Entry, LRU order
        self.captures = 0

    def reset(self):
        self.entries.clear()

    # ------------------------------------------------------------------ entry selection
    def _select(self, skey, length):
        best = None
        for (k, bounds), ent in self.entries.items():
            if k == skey and ent.plan.fits(length):
                if best is None or bounds < best[0]:
                    best = (bounds, ent)
        if best is not None:
            self.entries.move_to_end((skey, best[0]))
            return best[1]
        return None

    def _bounds_for_miss(self, skey, length):
        """First capture of a shape: tight bounds; any later miss: the generic plan that fits every list."""
        s = sum(length)
        if any(k == skey for k, _ in self.entries):
            return (s, s)
        return (len(length), max(length))

    def __call__(self, x, pos_mask, length, device=None, bounds=None):
        """`length`: the sequence structure handed to the plan (persons per image).  `bounds` = (images, persons per
        image) fixes the launch bounds instead of the tight-then-generic policy (sharded.ShardedForward)."""
        device = torch.device(device) if device is not None else x.device
        length = [int(n) for n in length]
        skey = (tuple(x.shape), tuple(pos_mask.shape), device.index)
        ent = self.entries.get((skey, tuple(bounds))) if bounds is not None else self._select(skey, length)
        if ent is None:
            bounds = tuple(bounds) if bounds is not None else self._bounds_for_miss(skey, length)
            while len(self.entries) >= self.max_entries:
                self.entries.popitem(last=False)
            ent = self._capture(x, pos_mask, length, device, bounds)
            self.entries[(skey, bounds)] = ent
        main = torch.cuda.current_stream(device)
        ent.plan.update(length)
        host = not (x.is_cuda and pos_mask.is_cuda)
        if host:
            cs = ent.copy_stream
            with torch.cuda.stream(cs):
                cs.wait_event(ent.x_stage_free)     # the previous call has drained stage_x (at its very start)
                ent.stage_x.copy_(x, non_blocking=True)
                ent.x_ready.record(cs)
                cs.wait_event(ent.m_stage_free)     # ... and stage_m (where its graphs are cut)
                ent.stage_m.copy_(pos_mask, non_blocking=True)
                ent.mask_ready.record(cs)
            main.wait_event(ent.x_ready)
            ent.sx.copy_(ent.stage_x, non_blocking=True)
            ent.x_stage_free.record(main)
        else:
            ent.sx.copy_(x, non_blocking=True)
        mask_done = False
        for pre, graph in ent.segments:
            if pre == "mask":
                self._hand_over_mask(ent, pos_mask, host, main)
                mask_done = True
            elif pre is not None:
                pre()
            graph.replay()
        if not mask_done:                           # the sequence never reads pos_mask: keep the staging protocol going
            self._hand_over_mask(ent, pos_mask, host, main)
        return _clone_tree(ent.out)

    @staticmethod
    def _hand_over_mask(ent, pos_mask, host, main):
        if host:
            main.wait_event(ent.mask_ready)
            ent.sm.copy_(ent.stage_m, non_blocking=True)
            ent.m_stage_free.record(main)
        else:
            ent.sm.copy_(pos_mask, non_blocking=True)

    def _capture(self, x, pos_mask, length, device, bounds):
        ent = _Entry()
        ent.plan = SeqPlan(device, sum(length), bounds[0], bounds[1])
        ent.plan.length = list(length)
        ent.sx = x.to(device, dtype=torch.float32, copy=True).contiguous()
        ent.sm = pos_mask.to(device, dtype=torch.float32, copy=True).contiguous()
        ent.stage_x = torch.empty_like(ent.sx)
        ent.stage_m = torch.empty_like(ent.sm)
        main = torch.cuda.current_stream(device)
        side = torch.cuda.Stream(device=device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self.eager_fn(ent.sx, ent.sm, ent.plan, EagerHooks())      # warm-up outside capture (creates the plan's buffers)
        main.wait_stream(side)
        torch.cuda.synchronize(device)
        hooks = _CaptureHooks(torch.cuda.graph_pool_handle())
        try:
            out = self.eager_fn(ent.sx, ent.sm, ent.plan, hooks)
        finally:
            hooks._close()
        self.captures += 1
        ent.segments, ent.out = hooks.segments, out
        ent.copy_stream = torch.cuda.Stream(device=device)
        ent.x_ready = torch.cuda.Event()
        ent.mask_ready = torch.cuda.Event()
        ent.x_stage_free = torch.cuda.Event()
        ent.m_stage_free = torch.cuda.Event()
        ent.x_stage_free.record(main)
        ent.m_stage_free.record(main)
        return ent


def _clone_tree(out):
    if isinstance(out, dict):
        return {k: v.clone() for k, v in out.items()}
    return out.clone()


class ExactPlans:
    """Bounded LRU of exact SeqPlans for the eager (non-graph) path."""

    def __init__(self, device, capacity=64):
        self.device, self.capacity = device, capacity
        self.plans = collections.OrderedDict()

    def get(self, length):
        key = tuple(length)
        plan = self.plans.get(key)
        if plan is None:
            while len(self.plans) >= self.capacity:
                self.plans.popitem(last=False)
            plan = SeqPlan.exact(length, self.device)
            self.plans[key] = plan
        else:
            self.plans.move_to_end(key)
        return plan


class HostPipeline:
    """Pipelined host-to-host inference around a drop-in module: `submit(x, pos_mask, length)` enqueues the upload, the
    forward and the download of the heatmaps into a ring of pinned host buffers and returns a ticket at once;
    `ticket.result()` waits for that batch only.  With depth >= 2 the upload of batch i+1 and the download of batch
    i-1 run on their own streams under the kernels of batch i -- the loop the reference's `validate`
    (lib/core/function.py:126-191) would need to keep a B200 busy:

        pipe = HostPipeline(model, depth=2)
        prev = None
        for x, pos_mask, ..., meta in loader:
            t = pipe.submit(x, pos_mask, meta['length'].tolist())
            if prev is not None: consume(prev.result())      # heatmaps of the previous batch, host memory
            prev = t
    """

    class Ticket:
        def __init__(self, event, host):
            self.event, self.host = event, host

        def result(self):
            self.event.synchronize()
            return self.host

    def __init__(self, model, depth=2):
        self.model, self.depth = model, int(depth)
        self.slots = [None] * self.depth
        self.i = 0
        self.d2h = None

    def submit(self, x, pos_mask, length):
        out = self.model(x, pos_mask, length)
        tensors = out if isinstance(out, dict) else {"out": out}
        dev = next(iter(tensors.values())).device
        if self.d2h is None:
            self.d2h = torch.cuda.Stream(device=dev)
        slot = self.slots[self.i % self.depth]
        if slot is None or any(slot[1][k].shape != v.shape for k, v in tensors.items()):
            host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in tensors.items()}
            slot = (torch.cuda.Event(), host)
            self.slots[self.i % self.depth] = slot
        else:
            slot[0].synchronize()          # the ticket that used this slot `depth` submits ago must be done
        self.i += 1
        ev, host = slot
        self.d2h.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.d2h):
            for k, v in tensors.items():
                host[k].copy_(v, non_blocking=True)
                v.record_stream(self.d2h)
            ev.record(self.d2h)
        return HostPipeline.Ticket(ev, host if isinstance(out, dict) else host["out"])
