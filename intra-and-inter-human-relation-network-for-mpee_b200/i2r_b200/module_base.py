"""Shared plumbing of the drop-in modules (`models.interformer_pureMulti`, `models.interformer`,
`models.interformer_2stage`): argument checks of `forward(x, pos_mask, length)` (lib/core/function.py:135), lazy
weight packing, CUDA-graph dispatch and invalidation of the packed device program when weights change.

The reference's workflows change weights in several ways the wrapper cannot intercept by overriding its own
`load_state_dict`: `model.singleformer.load_state_dict(...)` (interformer.py:147-155 loads the first stage that
way), `init_weights` on a sub-module, in-place edits of `p.data`.  All of them bump the autograd version counter of
the touched tensors, so every forward compares the sum of `_version` over all parameters and buffers with the value
recorded when the device program was packed and re-packs on a mismatch (~60 us of host time for the largest model,
hidden behind the asynchronous launches; `watch_weights = False` switches it off).
"""
import os

import torch
import torch.nn as nn

from . import capi
from .engine import ExactPlans, GraphedForward, SeqPlan


class DevicePathModule(nn.Module):
    flavor = "model"

    def _init_device_path(self, runner_factory):
        self._program = None
        self._graphs = GraphedForward(self._eager)
        self._exact = None
        self._weights_tag = None
        self._watched = None
        self.watch_weights = True
        self.use_cuda_graph = os.environ.get("I2R_CUDA_GRAPH", "1") != "0"
        self.check_impl = False     # tests: route implicit GEMMs through the scalar check kernel
        self._runner_factory = runner_factory

    # ------------------------------------------------------------------ weights -> device program
    def _weights_tag_now(self):
        if self._watched is None:
            self._watched = list(self.parameters()) + list(self.buffers())
        return sum(t._version for t in self._watched)

    def _program_ready(self, prog):
        self._program = prog
        self._graphs.reset()
        self._exact = ExactPlans(prog.device)
        self._watched = None
        self._weights_tag = self._weights_tag_now()

    def invalidate(self):
        """Forget the packed device program (and its captured graphs); the next forward packs the weights again."""
        self._program = None
        self._watched = None
        if getattr(self, "_graphs", None) is not None:
            self._graphs.reset()

    def load_state_dict(self, *a, **kw):
        out = super().load_state_dict(*a, **kw)
        self.invalidate()
        return out

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self.invalidate()
        return out

    def _ensure_program(self, dev):
        stale = self._program is None or self._program.device != dev
        if not stale and self.watch_weights and self._weights_tag_now() != self._weights_tag:
            stale = True
        if stale:
            self.prepare(dev)

    def _plan(self, length_or_plan):
        if isinstance(length_or_plan, SeqPlan):
            return length_or_plan
        if self._exact is None:
            self._exact = ExactPlans(self._program.device)
        return self._exact.get([int(n) for n in length_or_plan])

    # ------------------------------------------------------------------ visualisation hooks
    def _hook_targets(self):
        """(token-map layer or None, inter-human encoder holder): the sub-modules the reference's visualize.py:164-175
        registers forward hooks on (`model.reduce`, `model.global_encoder.layers[i].self_attn`)."""
        enc = getattr(self, "global_encoder", None) or getattr(self, "multi_global_encoder", None)
        return getattr(self, "reduce", None), enc

    def _hooked(self):
        red, enc = self._hook_targets()
        if red is not None and red._forward_hooks:
            return True
        return enc is not None and any(layer.self_attn._forward_hooks for layer in enc.layers)

    def _fire_hooks(self, module, inputs, output):
        for hook in list(module._forward_hooks.values()):
            hook(module, inputs, output)

    def _attn_tap(self, length, tokens):
        """The forward never materialises attention weights (streaming softmax in TMEM); when a hook asks for them they
        are rebuilt here, per image, from the projected queries and keys of the layer: [bs, L, L] fp32 with L =
        max(length) * tokens, zero for padded persons -- what nn.MultiheadAttention returns as `output[1]`."""
        red, enc = self._hook_targets()
        lmax = max(length) * tokens

        def tap(layer, q, k, scale):
            mod = enc.layers[layer].self_attn
            if not mod._forward_hooks:
                return
            w = torch.zeros((len(length), lmax, lmax), dtype=torch.float32, device=q.device)
            at = 0
            for b, n in enumerate(length):
                t = n * tokens
                w[b, :t, :t] = torch.softmax((q[at:at + t] * scale) @ k[at:at + t].t(), dim=-1)
                at += t
            self._fire_hooks(mod, (), (None, w))
        return tap

    # ------------------------------------------------------------------ forward
    def _eager(self, x, pos_mask, length, hooks=None):
        """The launch sequence of one forward.  `length`: persons per image, or an engine.SeqPlan (graph capture:
        offsets are device data).  `hooks.mask_needed()` marks the first use of `pos_mask` (engine.GraphedForward cuts
        its graphs there so the mask upload overlaps everything before it)."""
        p = self._program
        r = p.runner
        plan = self._plan(length)
        feat, heat_single, tok = self._stage_tokens(p, r, x)
        s, th, tw, d = tok.shape
        tap = None
        if hooks is None and self._hooked():      # eager calls only (forward disables graph replay while hooks exist)
            red, _ = self._hook_targets()
            if red is not None and red._forward_hooks:
                t32 = tok.float()
                if getattr(p, "split", False):
                    t32 = t32[..., :d // 2] + t32[..., d // 2:]
                self._fire_hooks(red, (), t32.permute(0, 3, 1, 2).contiguous())
            tap = self._attn_tap(plan.length, th * tw)
        pos = None
        if p.mask_embed is not None:
            if hooks is not None:
                hooks.mask_needed()
            pos = self._stage_pos(p, r, pos_mask, (th, tw)).view(s * th * tw, d)
        y = p.encoder.run(r, tok.view(s * th * tw, d), pos, plan.cu(th * tw), plan.max_seqlen(th * tw), attn_tap=tap)
        return self._stage_head(p, r, y.view(s, th, tw, d), feat, heat_single)

    def _device(self):
        return next(self.parameters()).device

    def forward(self, x, pos_mask, length):
        length = [int(n) for n in length]
        if sum(length) != x.shape[0] or x.shape[0] != pos_mask.shape[0]:
            raise ValueError("sum(length)=%d must equal the number of crops %d" % (sum(length), x.shape[0]))
        if min(length) < 1:
            raise ValueError("every image needs at least one person crop")
        dev = self._device()
        if dev.type != "cuda":
            raise capi.I2RError("%s forward runs on a CUDA (sm_100a) device only; move the module with .cuda() -- "
                                "there is no CPU fallback" % self.flavor)
        with torch.cuda.device(dev):
            self._ensure_program(dev)
            with torch.no_grad():
                if self.use_cuda_graph and not self._hooked():      # host tensors go through the engine's staging buffers
                    if x.dtype != torch.float32 or pos_mask.dtype != torch.float32:
                        x, pos_mask = x.float(), pos_mask.float()
                    return self._graphs(x, pos_mask, length, device=dev)
                x = x.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
                pos_mask = pos_mask.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
                return self._eager(x, pos_mask, length)

    def forward_flip(self, x, pos_mask, length, flip_pairs):
        """The reference's flip test (lib/core/function.py:142-162) as ONE forward: the crops and their mirror images
        go through the network as one batch of 2*S crops (the mirrored crops of an image form an image of their own, so
        the inter-human stage sees exactly what the reference's second forward sees), then
        `(out + flip_back(out_flipped, flip_pairs)) * 0.5` (lib/utils/transforms.py:16-30) is one kernel on the device --
        no numpy flips, no second launch sequence, no D2H / H2D round trip.  Returns what forward returns."""
        from . import postproc
        length = [int(n) for n in length]
        dev = self._device()
        if dev.type != "cuda":
            raise capi.I2RError("%s forward_flip runs on a CUDA (sm_100a) device only" % self.flavor)
        with torch.cuda.device(dev), torch.no_grad():
            x = x.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
            pos_mask = pos_mask.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
            s = x.shape[0]
            x2 = torch.cat([x, postproc.hflip(x)], 0)
            m2 = torch.cat([pos_mask, postproc.hflip(pos_mask)], 0)
            out = self.forward(x2, m2, length + length)

            def merge(o):
                return postproc.flip_merge(o[:s], o[s:], flip_pairs)
            return {k: merge(v) for k, v in out.items()} if isinstance(out, dict) else merge(out)
