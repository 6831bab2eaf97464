"""CUDA-graph cache around an eager kernel-launch sequence.

The forward is a fixed sequence of ~100-150 kernel launches whose shapes depend only on
(number of crops, persons per image, input size).  The first call for a shape runs the sequence
eagerly once (warm-up: lazy module loading, allocator growth), then captures it into CUDA graphs
with static input/output buffers; later calls copy the inputs into the static buffers and replay.

The sequence is captured as TWO graphs, cut at the first use of `pos_mask` (the eager function calls
the `mask_needed` hook there): the image tensor is uploaded on the launching stream, the box masks on
a second stream, and only the second graph waits for them -- so with host inputs (the reference's
tools/test.py hands the module CPU tensors, tools/test.py:118) the mask upload overlaps the backbone.
"""
import torch


class _Entry:
    __slots__ = ("sx", "sm", "g1", "g2", "out", "copy_stream", "mask_ready", "main_done")


class GraphedForward:
    def __init__(self, eager_fn, max_entries=8):
        self.eager_fn = eager_fn          # eager_fn(x, pos_mask, length, mask_needed=None)
        self.max_entries = max_entries
        self.entries = {}

    def reset(self):
        self.entries.clear()

    @staticmethod
    def seq_offsets(length, tokens_per_person, device):
        off = [0]
        for n in length:
            off.append(off[-1] + n * tokens_per_person)
        return torch.tensor(off, dtype=torch.int32).to(device)

    def __call__(self, x, pos_mask, length, device=None):
        device = torch.device(device) if device is not None else x.device
        key = (tuple(x.shape), tuple(pos_mask.shape), tuple(length), device.index)
        ent = self.entries.get(key)
        if ent is None:
            if len(self.entries) >= self.max_entries:
                self.entries.pop(next(iter(self.entries)))
            ent = self._capture(x, pos_mask, length, device)
            self.entries[key] = ent
        main = torch.cuda.current_stream(device)
        if ent.g2 is None:
            ent.sx.copy_(x, non_blocking=True)
            ent.sm.copy_(pos_mask, non_blocking=True)
            ent.g1.replay()
            return _clone_tree(ent.out)
        # masks on the copy stream (after the previous replay has finished reading the static buffer)
        ent.copy_stream.wait_event(ent.main_done)
        with torch.cuda.stream(ent.copy_stream):
            ent.sm.copy_(pos_mask, non_blocking=True)
            ent.mask_ready.record(ent.copy_stream)
        ent.sx.copy_(x, non_blocking=True)
        ent.g1.replay()
        main.wait_event(ent.mask_ready)
        ent.g2.replay()
        ent.main_done.record(main)
        return _clone_tree(ent.out)

    def _capture(self, x, pos_mask, length, device):
        ent = _Entry()
        ent.sx = x.to(device, dtype=torch.float32, copy=True).contiguous()
        ent.sm = pos_mask.to(device, dtype=torch.float32, copy=True).contiguous()
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            self.eager_fn(ent.sx, ent.sm, length)          # warm-up outside capture
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        pool = torch.cuda.graph_pool_handle()
        g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        ctx = [torch.cuda.graph(g1, pool=pool), None]

        def mask_needed():
            # cut: everything recorded so far needs the images only
            if ctx[1] is None:
                ctx[0].__exit__(None, None, None)
                ctx[1] = torch.cuda.graph(g2, pool=pool)
                ctx[1].__enter__()
        ctx[0].__enter__()
        try:
            out = self.eager_fn(ent.sx, ent.sm, length, mask_needed)
        except BaseException:
            (ctx[1] or ctx[0]).__exit__(None, None, None)
            raise
        (ctx[1] or ctx[0]).__exit__(None, None, None)
        ent.g1, ent.g2, ent.out = g1, (g2 if ctx[1] is not None else None), out
        ent.copy_stream = torch.cuda.Stream(device=device)
        ent.mask_ready = torch.cuda.Event()
        ent.main_done = torch.cuda.Event()
        ent.main_done.record(torch.cuda.current_stream(device))
        return ent


def _clone_tree(out):
    if isinstance(out, dict):
        return {k: v.clone() for k, v in out.items()}
    return out.clone()
