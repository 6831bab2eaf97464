"""Crop-sharded forward with ONE all-gather of the pooled token maps (SURVEY.md 8e partitioning (2); north_star:
"the batch of person crops shards naturally across the 8 GPUs, with a single NCCL all-gather over NVLink of per-image
instance tokens for the inter-human stage").

The reference forward has no collective (all persons of an image sit in one batch on one device,
lib/dataset/collater.py:19-24); its only coupling between crops is the inter-human encoder, and only among crops of
the same image (lib/models/attention.py:131-137, lib/models/interformer.py:290-306).  Here:

    every rank                     :  gets the whole batch (x, pos_mask, length), as every rank of the reference's
                                      tools/ddp_test.py does (:156-172), and takes its contiguous slice of the S crops
    per-crop stages (local crops)  :  backbone (+ intra-human encoder) -> pooled token map [S_r, h*w, d]  (+ mask embedding)
    ONE all_gather_into_tensor     :  token maps (and position embeddings) of all crops, <= 12.9 MB fp32-equivalent at C5
    inter-human encoder            :  on the window of crops around the local slice that holds every image touching the
                                      slice whole (sharding.CropShardLayout); images straddling a slice boundary are
                                      encoded on both sides, nothing is sent back
    upsample + residual + head     :  local crops only -> local heatmaps (optionally all-gathered into batch order)

Everything that depends on `length` is device data (engine.SeqPlan), so one set of CUDA-graph segments per batch size
serves every persons-per-image list; the collective runs eagerly between two graph segments on the launching stream.
"""
import torch
import torch.nn as nn

from . import capi
from .engine import EagerHooks, GraphedForward, SeqPlan
from .sharding import CropShardLayout


class ShardedForward(nn.Module):
    """Wraps a drop-in module (`models.interformer_pureMulti` / `interformer` / `interformer_2stage` instance).

    forward(x, pos_mask, length) -> heatmaps of THIS rank's crops `[c0, c1)` (`crop_range(S)`), or of the whole batch
    in order when gather_output=True.  `persons_bound`: the largest number of persons per image the caller will send
    (DATASET.MAX_PATCH, or the batch layout of a benchmark); None = no bound (the window is then the whole batch)."""

    def __init__(self, model, process_group=None, persons_bound=None, gather_output=False, world_size=None, rank=None):
        super().__init__()
        import torch.distributed as dist
        self.model = model
        self.group = process_group
        if world_size is None:
            world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
            rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.world, self.rank = int(world_size), int(rank)
        self.persons_bound = persons_bound
        self.gather_output = gather_output
        self.use_cuda_graph = getattr(model, "use_cuda_graph", True)
        self._graphs = GraphedForward(self._eager)
        self._layouts = {}
        self._lay = None
        self.bytes_gathered = 0          # per call, this rank's receive size (reported by bench.py)

    def layout(self, total_crops):
        key = int(total_crops)
        if key not in self._layouts:
            pb = self.persons_bound if self.persons_bound is not None else key
            self._layouts[key] = CropShardLayout(key, self.world, self.rank, pb)
        return self._layouts[key]

    def crop_range(self, total_crops):
        lay = self.layout(total_crops)
        return lay.c0, lay.c1

    # ------------------------------------------------------------------ the launch sequence of one rank
    def _all_gather(self, out, inp):
        import torch.distributed as dist
        if self.world == 1:
            out.copy_(inp.view_as(out))
        else:
            dist.all_gather_into_tensor(out, inp, group=self.group)

    def _eager(self, xl, pml, plan, hooks=None):
        hooks = hooks if hooks is not None else EagerHooks()
        m, lay = self.model, self._lay
        p = m._program
        r = p.runner
        feat, heat_single, tok = m._stage_tokens(p, r, xl)              # local crops: [s_max, th, tw, dw]
        s, th, tw, dw = tok.shape
        t = th * tw
        parts = 2 if p.mask_embed is not None else 1
        send = torch.empty((s, parts, t * dw), dtype=tok.dtype, device=tok.device)
        send[:, 0].copy_(tok.view(s, t * dw))
        if parts == 2:
            hooks.mask_needed()
            send[:, 1].copy_(m._stage_pos(p, r, pml, (th, tw)).view(s, t * dw))
        gathered = torch.empty((self.world * s, parts, t * dw), dtype=tok.dtype, device=tok.device)
        self.bytes_gathered = gathered.numel() * gathered.element_size()
        hooks.between(lambda: self._all_gather(gathered, send))          # the only collective of the forward
        win = gathered[lay.w0: lay.w0 + lay.wn]                          # global crop order; constant offset
        if parts == 1:
            tokw, posw = win.view(lay.wn * t, dw), None
        else:
            both = torch.empty((2, lay.wn, t * dw), dtype=tok.dtype, device=tok.device)
            both[0].copy_(win[:, 0])
            both[1].copy_(win[:, 1])
            tokw, posw = both[0].view(lay.wn * t, dw), both[1].view(lay.wn * t, dw)
        y = p.encoder.run(r, tokw, posw, plan.cu(t), plan.max_seqlen(t)).view(lay.wn, th, tw, dw)
        y_loc = y[lay.local_offset: lay.local_offset + s]
        return m._stage_head(p, r, y_loc, feat, heat_single)

    # ------------------------------------------------------------------ forward
    @staticmethod
    def _local(t, lay):
        loc = t[lay.c0: lay.c1]
        if loc.shape[0] < lay.s_max:      # tail rank of an uneven split: zero crops keep the shapes static
            pad = torch.zeros((lay.s_max - loc.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            loc = torch.cat([loc, pad], 0)
        return loc

    def _run(self, x, pos_mask, length, dev):
        lay = self.layout(x.shape[0])
        self._lay = lay
        xl, pml = self._local(x, lay), self._local(pos_mask, lay)
        wl = lay.window_lengths(length)
        if self.use_cuda_graph and dev.type == "cuda":
            out = self._graphs(xl.float(), pml.float(), wl, device=dev, bounds=(lay.wn, lay.persons_bound))
        else:
            plan = SeqPlan(dev, lay.wn, lay.wn, lay.persons_bound)
            plan.update(wl)
            xl = xl.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
            pml = pml.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
            out = self._eager(xl, pml, plan)
        n = lay.c1 - lay.c0
        if not self.gather_output:
            return {k: v[:n] for k, v in out.items()} if isinstance(out, dict) else out[:n]

        def gather(v):
            full = torch.empty((self.world * lay.s_max,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
            self._all_gather(full, v.contiguous())
            return full[: lay.total]
        return {k: gather(v) for k, v in out.items()} if isinstance(out, dict) else gather(out)

    def forward(self, x, pos_mask, length):
        length = [int(n) for n in length]
        if sum(length) != x.shape[0] or x.shape[0] != pos_mask.shape[0]:
            raise ValueError("sum(length)=%d must equal the number of crops %d" % (sum(length), x.shape[0]))
        if min(length) < 1:
            raise ValueError("every image needs at least one person crop")
        dev = self.model._device()
        if dev.type != "cuda":
            raise capi.I2RError("the sharded forward runs on CUDA (sm_100a) devices only -- there is no CPU fallback")
        with torch.cuda.device(dev):
            if self.model._program is None or self.model._program.device != dev or (
                    self.model.watch_weights and self.model._weights_tag_now() != self.model._weights_tag):
                self.model.prepare(dev)
                self._graphs.reset()
            with torch.no_grad():
                return self._run(x, pos_mask, length, dev)
