"""CUDA-graph cache around an eager kernel-launch sequence.

The forward is a fixed sequence of ~150 kernel launches whose shapes depend only on
(number of crops, persons per image, input size).  The first call for a shape runs the sequence
eagerly once (warm-up: lazy module loading, allocator growth), then captures it into a CUDA graph
with static input/output buffers; later calls copy the inputs into the static buffers and replay.
"""
import torch


class GraphedForward:
    def __init__(self, eager_fn, max_entries=8):
        self.eager_fn = eager_fn
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

    def __call__(self, x, pos_mask, length):
        key = (tuple(x.shape), tuple(pos_mask.shape), tuple(length), x.device.index)
        ent = self.entries.get(key)
        if ent is None:
            if len(self.entries) >= self.max_entries:
                self.entries.pop(next(iter(self.entries)))
            ent = self._capture(x, pos_mask, length)
            self.entries[key] = ent
        sx, sm, graph, out = ent
        sx.copy_(x, non_blocking=True)
        sm.copy_(pos_mask, non_blocking=True)
        graph.replay()
        return _clone_tree(out)

    def _capture(self, x, pos_mask, length):
        sx, sm = x.clone(), pos_mask.clone()
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):
            self.eager_fn(sx, sm, length)          # warm-up outside capture
        torch.cuda.current_stream(x.device).wait_stream(side)
        torch.cuda.synchronize(x.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.eager_fn(sx, sm, length)
        return sx, sm, graph, out


def _clone_tree(out):
    if isinstance(out, dict):
        return {k: v.clone() for k, v in out.items()}
    return out.clone()
