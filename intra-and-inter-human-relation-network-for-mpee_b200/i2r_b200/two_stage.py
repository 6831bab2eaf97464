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
from .module_base import DevicePathModule
from .hrnet_w48 import conv_bn_layer
from .modules import DeconvProgram, EncoderParams, make_deconv_stack
from .ops import Runner, channel_padding, split_precision
from .position import MaskEmbedParams, build_mask_embed_program


class TwoStageInterFormer(DevicePathModule):
    def __init__(self, cfg, singleformer, flavor):
        super().__init__()
        assert flavor in ("interformer", "interformer_2stage")
        m, extra = cfg.MODEL, cfg.MODEL.EXTRA
        self.flavor = flavor
        # cfg.MODEL.SINGLEFORMER empty: the stand-alone HRNet produces the token maps directly (interformer.py:143,
        # :291-292: no first-stage heatmaps, no max-pool, no residual), its parameters live under `backbone.body.*`
        self.have_singleformer = singleformer is not None
        if self.have_singleformer:
            self.singleformer = singleformer
        else:
            import models.backbone
            self.backbone = models.backbone.build_backbone(cfg)
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
        if flavor == "interformer" and m.get("NORMALIZE_BEFORE", False):
            # lib/models/attention.py:1040 -- only this flavor's encoder honours the flag; silently running post-norm
            # would return wrong heatmaps
            raise NotImplementedError("NORMALIZE_BEFORE=True (pre-norm encoder; every shipped config uses post-norm)")
        if m.get("DOMAIN_TRANS", False):
            raise NotImplementedError("DOMAIN_TRANS=True (no shipped config enables it)")
        self.multi_position_embedding = MaskEmbedParams(self.trans_size, d_model, mode=self.multi_position_mode,
                                                        vec_dim=m.MULTI_POS_EMBEDDING_DIM)
        self.multi_global_encoder = EncoderParams(d_model, m.N_HEAD, m.DIM_FEEDFORWARD, m.ENCODER_MULTI_LAYERS)
        self.deconv_with_bias = bool(extra.DECONV_WITH_BIAS)
        # number of x2 upsampling steps from the token map back to the heatmap
        self.up_steps = int(math.log(self.heatmap_size[0] // self.trans_size[1], 2))
        if self.upsample_type == "upconv":      # interformer.py:25-64
            holder = nn.Module()
            scale = self.heatmap_size[0] // self.trans_size[1]
            holder.fuse_layers = nn.Sequential(nn.Conv2d(d_model, d_model, 1, 1, 0, bias=False), nn.BatchNorm2d(d_model),
                                               nn.Upsample(scale_factor=scale, mode="nearest"))
            holder.double_conv = nn.Sequential(
                nn.Conv2d(d_model, d_model, 3, padding=1, bias=False), nn.BatchNorm2d(d_model), nn.ReLU(inplace=True),
                nn.Conv2d(d_model, d_model, 3, padding=1, bias=False), nn.BatchNorm2d(d_model), nn.ReLU(inplace=True))
            self.upsample_layer = holder
            self._upconv_shift = int(math.log(scale, 2))
            if 2 ** self._upconv_shift != scale:
                raise NotImplementedError("UpConv scale factor %d is not a power of two" % scale)
            self._deconv_keys = []
        elif self.upsample_type == "multiplex":
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
        self._init_device_path(Runner)

    @property
    def returns_dict(self):
        return self.inter_supervision and self.have_singleformer and not self.singleformer_fix

    # ------------------------------------------------------------------ weights -> device program
    def prepare(self, device=None):
        device = torch.device(device) if device is not None else self.final_layer.weight.device
        sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        c = self._cfg
        if c["final_k"] != 1:
            raise NotImplementedError("FINAL_CONV_KERNEL=3")
        if self.use_multi_pos and self.multi_position_mode not in ("conv", "res"):
            raise NotImplementedError("MULTI_POS_EMBEDDING=%r with USE_MULTI_POS (kernels exist for 'conv' and 'res')" %
                                      self.multi_position_mode)
        prog = type("Program", (), {})()
        prog.device = device
        prog.runner = self._runner_factory(device, 1 if self.check_impl else 0)
        # the whole model shares the first stage's precision mode (split-operand unless overridden)
        first_stage = self.singleformer if self.have_singleformer else self.backbone.body
        prog.split = getattr(first_stage, "precision", "fp16") == "split"
        prog.runner.split = prog.split
        prog.first = first_stage.build_program(device)
        with split_precision(prog.split), channel_padding(16 if c["d_model"] % 16 else 0):
            self._build_second_stage(prog, sd, c, device)
        self._program_ready(prog)
        return self

    def _build_second_stage(self, prog, sd, c, device):
        prog.mask_embed = build_mask_embed_program(self.multi_position_mode, sd, "multi_position_embedding",
                                                   device) if self.use_multi_pos else None
        prog.encoder = EncoderProgram(sd, "multi_global_encoder", c["layers"], c["d_model"], c["nhead"], device)
        cache = {}

        def deconv(key):
            if key not in cache:
                cache[key] = [DeconvProgram(sd, "%s.%d" % (key, 3 * i), "%s.%d" % (key, 3 * i + 1), device)
                              for i in range(c["num_deconv"])]
            return cache[key]
        prog.upsample = [deconv(k) for k in self._deconv_keys]
        prog.upconv = None
        if self.upsample_type == "upconv":
            u = "upsample_layer."
            prog.upconv = (conv_bn_layer(sd, u + "fuse_layers.0", u + "fuse_layers.1", device=device),
                           conv_bn_layer(sd, u + "double_conv.0", u + "double_conv.1", relu=True, device=device),
                           conv_bn_layer(sd, u + "double_conv.3", u + "double_conv.4", relu=True, device=device))
        prog.head = conv_bn_layer(sd, "final_layer", None, device=device)

    # ------------------------------------------------------------------ forward
    # the four stages DevicePathModule._eager (and sharded.ShardedForward) compose
    def _stage_tokens(self, p, r, x):
        if not self.have_singleformer:
            feats = p.first.backbone.run(r, x)
            return None, None, r.conv(p.first.reduce, feats[-1])            # token map at TRANS_SIZE already
        feat, heat_single = p.first.run(r, x)                                # [S,h,w,d] fp16, [S,K,h,w] fp32
        tok = feat
        for _ in range(int(math.log(feat.shape[2] // self.trans_size[-1], 2))):   # interformer.py:260-264
            tok = r.maxpool(tok)
        return feat, heat_single, tok

    def _stage_pos(self, p, r, pos_mask, hw):
        return None if p.mask_embed is None else p.mask_embed.run(r, pos_mask, hw)

    def _stage_head(self, p, r, y, feat, heat_single):
        for stack in p.upsample:
            for dc in stack:
                y = dc.run(r, y)
        if p.upconv is not None:
            # UpConv: 1x1 + BN at the token resolution (commutes with the nearest up-sampling that follows it), then the
            # first 3x3 reads the map through a >> shift gather, the second runs at the heatmap resolution
            y = r.conv(p.upconv[0], y)
            y = r.conv(p.upconv[1], y, in_shift=self._upconv_shift)
            y = r.conv(p.upconv[2], y)
        if feat is not None:
            y = r.add(feat, y)                                               # single_res + x
        heat_multi = r.conv(p.head, y, out_mode="nchw32")
        if self.returns_dict:
            return {"single": heat_single, "multi": heat_multi}
        return heat_multi


def build(cfg, is_train, flavor, models_pkg):
    """`get_pose_net` body of both reference modules: the first stage is resolved by name exactly as
    interformer.py:139 / interformer_2stage.py:428 do (`eval('models.' + cfg.MODEL.SINGLEFORMER + '.get_pose_net')`)."""
    name = cfg.MODEL.SINGLEFORMER
    if not name:
        return TwoStageInterFormer(cfg, None, flavor)
    factory = getattr(getattr(models_pkg, name), "get_pose_net")
    single = factory(cfg, is_train, cfg.MODEL.SINGLE_MODEL, cfg.MODEL.END2END)
    return TwoStageInterFormer(cfg, single, flavor)
