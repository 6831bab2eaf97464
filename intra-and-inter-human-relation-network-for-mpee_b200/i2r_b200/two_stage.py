"""Two-stage I2R-Net wrapper shared by `models.interformer` and `models.interformer_2stage`:

    first stage (per crop)  ->  max-pool to TRANS_SIZE  ->  inter-human encoder over the ragged N*h*w token
    sequence of every image (+ conv position embedding of the person box masks)  ->  deconv upsample back to the
    heatmap size  ->  + first-stage feature  ->  1x1 heatmap head.

Reference: lib/models/interformer.py:129-323 and lib/models/interformer_2stage.py:208-423.  The two reference
modules differ only in parameter names of the upsample path and in which encoder class they instantiate (same
arithmetic); `flavor` selects the naming so the state_dict keys match either one.
"""
import math
import os

import torch
import torch.nn as nn

from . import capi
from .encoder import EncoderProgram
from .engine import GraphedForward
from .hrnet_w48 import conv_bn_layer
from .modules import DeconvProgram, EncoderParams, make_deconv_stack
from .ops import Runner, channel_padding, split_precision
from .position import MaskEmbedParams, MaskEmbedProgram


class TwoStageInterFormer(nn.Module):
    def __init__(self, cfg, singleformer, flavor):
        super().__init__()
        assert flavor in ("interformer", "interformer_2stage")
        m, extra = cfg.MODEL, cfg.MODEL.EXTRA
        self.flavor = flavor
        self.singleformer = singleformer
        self.singleformer_fix = bool(m.SINGLEFORMER_FIX)
        self.trans_size = list(m.TRANS_SIZE)
        self.heatmap_size = list(m.HEATMAP_SIZE)
        d_model = m.DIM_MODEL
        self.use_multi_pos = bool(m.USE_MULTI_POS)
        self.inter_supervision = bool(m.INTER_SUPERVISION)
        self.upsample_type = m.UPSAMPLE_TYPE
        self.multi_position_mode = m.MULTI_POS_EMBEDDING
        if m.get("ATTENTION_TYPE", "default") != "default":
            raise NotImplementedError("ATTENTION_TYPE=%r (every shipped config uses 'default')" % m.ATTENTION_TYPE)
        if m.get("DOMAIN_TRANS", False):
            raise NotImplementedError("DOMAIN_TRANS=True (no shipped config enables it)")
        self.multi_position_embedding = MaskEmbedParams(self.trans_size, d_model, mode=self.multi_position_mode,
                                                        vec_dim=m.MULTI_POS_EMBEDDING_DIM)
        self.multi_global_encoder = EncoderParams(d_model, m.N_HEAD, m.DIM_FEEDFORWARD, m.ENCODER_MULTI_LAYERS)
        self.deconv_with_bias = bool(extra.DECONV_WITH_BIAS)
        # number of x2 upsampling steps from the token map back to the heatmap
        self.up_steps = int(math.log(self.heatmap_size[0] // self.trans_size[1], 2))
        if self.upsample_type == "upconv":
            raise NotImplementedError("UPSAMPLE_TYPE='upconv' (no shipped config uses it)")
        if self.upsample_type == "multiplex":
            self.deconv_layers = make_deconv_stack(extra, self.deconv_with_bias)
            self._deconv_keys = ["deconv_layers"] * (self.up_steps if flavor == "interformer_2stage" else 2)
        elif self.upsample_type == "deconv":
            if flavor == "interformer":          # interformer.py:67-127 -- `upsample_layer.deconv_layers.<i>`
                holder = nn.Module()
                holder.deconv_layers = nn.ModuleList(make_deconv_stack(extra, self.deconv_with_bias)
                                                     for _ in range(self.up_steps))
                self.upsample_layer = holder
                self._deconv_keys = ["upsample_layer.deconv_layers.%d" % i for i in range(self.up_steps)]
            else:                                # interformer_2stage.py:244-260 -- deconv_layers1..3
                for i in (1, 2, 3):
                    setattr(self, "deconv_layers%d" % i, make_deconv_stack(extra, self.deconv_with_bias))
                self._deconv_keys = ["deconv_layers%d" % (i + 1) for i in range(self.up_steps)]
        else:
            raise ValueError("unknown UPSAMPLE_TYPE %r" % self.upsample_type)
        k = extra["FINAL_CONV_KERNEL"]
        self.final_layer = nn.Conv2d(d_model, m.NUM_JOINTS, k, 1, 1 if k == 3 else 0)
        self._cfg = dict(d_model=d_model, nhead=m.N_HEAD, layers=m.ENCODER_MULTI_LAYERS, final_k=k,
                         num_deconv=extra.NUM_DECONV_LAYERS)
        self._program = None
        self._graphs = GraphedForward(self._eager)
        self.use_cuda_graph = os.environ.get("I2R_CUDA_GRAPH", "1") != "0"
        self.check_impl = False
        self._runner_factory = Runner

    @property
    def returns_dict(self):
        return self.inter_supervision and not self.singleformer_fix

    # ------------------------------------------------------------------ weights -> device program
    def prepare(self, device=None):
        device = torch.device(device) if device is not None else self.final_layer.weight.device
        sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        c = self._cfg
        if c["final_k"] != 1:
            raise NotImplementedError("FINAL_CONV_KERNEL=3")
        if self.use_multi_pos and self.multi_position_mode != "conv":
            raise NotImplementedError("MULTI_POS_EMBEDDING=%r with USE_MULTI_POS (kernels exist for 'conv')" %
                                      self.multi_position_mode)
        prog = type("Program", (), {})()
        prog.device = device
        prog.runner = self._runner_factory(device, 1 if self.check_impl else 0)
        # the whole model shares the first stage's precision mode (split-operand unless overridden)
        prog.split = getattr(self.singleformer, "precision", "fp16") == "split"
        prog.runner.split = prog.split
        prog.first = self.singleformer.build_program(device)
        with split_precision(prog.split), channel_padding(16 if c["d_model"] % 16 else 0):
            self._build_second_stage(prog, sd, c, device)
        offsets = {}

        def seq_offsets(length, tokens_per_person):
            key = (tuple(length), tokens_per_person)
            if key not in offsets:
                offsets[key] = GraphedForward.seq_offsets(length, tokens_per_person, device)
            return offsets[key]
        prog.seq_offsets = seq_offsets
        self._program = prog
        self._graphs.reset()
        return self

    def _build_second_stage(self, prog, sd, c, device):
        prog.mask_embed = MaskEmbedProgram(sd, "multi_position_embedding", device) if self.use_multi_pos else None
        prog.encoder = EncoderProgram(sd, "multi_global_encoder", c["layers"], c["d_model"], c["nhead"], device)
        cache = {}

        def deconv(key):
            if key not in cache:
                cache[key] = [DeconvProgram(sd, "%s.%d" % (key, 3 * i), "%s.%d" % (key, 3 * i + 1), device)
                              for i in range(c["num_deconv"])]
            return cache[key]
        prog.upsample = [deconv(k) for k in self._deconv_keys]
        prog.head = conv_bn_layer(sd, "final_layer", None, device=device)

    def load_state_dict(self, *a, **kw):
        out = super().load_state_dict(*a, **kw)
        self._program = None
        return out

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._program = None
        return out

    # ------------------------------------------------------------------ forward
    def _eager(self, x, pos_mask, length, mask_needed=None):
        p = self._program
        r = p.runner
        feat, heat_single = self._run_first(p, r, x)                         # [S,h,w,d] fp16, [S,K,h,w] fp32
        tok = feat
        for _ in range(int(math.log(feat.shape[2] // self.trans_size[-1], 2))):   # interformer.py:260-264
            tok = r.maxpool(tok)
        s, th, tw, d = tok.shape
        pos = None
        if p.mask_embed is not None:
            if mask_needed is not None:
                mask_needed()         # engine.GraphedForward: the graph is cut here (mask upload overlaps what precedes)
            pos = p.mask_embed.run(r, pos_mask, (th, tw)).view(s * th * tw, d)
        cu = p.seq_offsets(length, th * tw)
        y = p.encoder.run(r, tok.view(s * th * tw, d), pos, cu, max(length) * th * tw).view(s, th, tw, d)
        for stack in p.upsample:
            for dc in stack:
                y = dc.run(r, y)
        y = r.add(feat, y)                                                   # single_res + x
        heat_multi = r.conv(p.head, y, out_mode="nchw32")
        if self.returns_dict:
            return {"single": heat_single, "multi": heat_multi}
        return heat_multi

    @staticmethod
    def _run_first(p, r, x):
        """First stage over all crops.  In split-operand mode the crops go through in groups: the persistent conv kernel
        has a latent fault (launch failure, timing dependent, clean under compute-sanitizer) when a CTA of a
        STREAMED-weight split-mode problem runs four or more tiles (first seen on the 64 -> 256 1x1 + residual of
        layer1 at >= 24 crops).  Per-crop stages are independent, so grouping is exact; the group
        size keeps the largest GEMM of the stage at <= 3 tiles per CTA."""
        s, _, h, w = x.shape
        if not p.split:
            return p.first.run(r, x)
        rows = max((h // 4) * (w // 4), ((h // 4 + 6) // 7 * 7) * ((w // 4 + 6) // 7 * 7))   # pixels / window rows per crop
        group = max(1, (3 * 148 * 128) // rows)
        if s <= group:
            return p.first.run(r, x)
        feats, heats = [], []
        for i in range(0, s, group):
            f, hm = p.first.run(r, x[i:i + group])
            feats.append(f)
            heats.append(hm)
        return torch.cat(feats, 0), torch.cat(heats, 0)

    def forward(self, x, pos_mask, length):
        length = [int(n) for n in length]
        if sum(length) != x.shape[0] or x.shape[0] != pos_mask.shape[0]:
            raise ValueError("sum(length)=%d must equal the number of crops %d" % (sum(length), x.shape[0]))
        if min(length) < 1:
            raise ValueError("every image needs at least one person crop")
        dev = self.final_layer.weight.device
        if dev.type != "cuda":
            raise capi.I2RError("%s forward runs on a CUDA (sm_100a) device only; move the module with .cuda() -- "
                                "there is no CPU fallback" % self.flavor)
        if self._program is None or self._program.device != dev:
            self.prepare(dev)
        with torch.no_grad():
            if self.use_cuda_graph:      # host tensors are uploaded straight into the graphs' static buffers
                if x.dtype != torch.float32 or pos_mask.dtype != torch.float32:
                    x, pos_mask = x.float(), pos_mask.float()
                return self._graphs(x, pos_mask, length, device=dev)
            x = x.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
            pos_mask = pos_mask.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
            return self._eager(x, pos_mask, length)


def build(cfg, is_train, flavor, models_pkg):
    """`get_pose_net` body of both reference modules: the first stage is resolved by name exactly as
    interformer.py:139 / interformer_2stage.py:428 do (`eval('models.' + cfg.MODEL.SINGLEFORMER + '.get_pose_net')`)."""
    name = cfg.MODEL.SINGLEFORMER
    if not name:
        raise NotImplementedError("MODEL.SINGLEFORMER is empty: the stand-alone HRNet path (lib/models/hrnet.py) is "
                                  "not referenced by any shipped config")
    factory = getattr(getattr(models_pkg, name), "get_pose_net")
    single = factory(cfg, is_train, cfg.MODEL.SINGLE_MODEL, cfg.MODEL.END2END)
    return TwoStageInterFormer(cfg, single, flavor)
