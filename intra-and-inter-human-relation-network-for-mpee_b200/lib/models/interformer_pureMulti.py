"""Vanilla I2R-Net ("interformer_pureMulti"): HRNet-W48-S -> reduce -> inter-human encoder over the
16x12 token maps of all persons of an image -> the same deconv applied twice -> 1x1 heatmap head.

Drop-in for the reference module of the same name (lib/models/interformer_pureMulti.py):
`get_pose_net(cfg, is_train)` returns an nn.Module with identical state_dict keys/shapes whose
`forward(x, pos_mask, length)` runs entirely in the sm_100a kernels of libi2r_sm100.so.
There is no torch/CPU forward here: calling it off-GPU raises.
"""
import logging
import math
import os

import torch
import torch.nn as nn

from i2r_b200 import capi
from i2r_b200.encoder import EncoderProgram
from i2r_b200.hrnet_w48 import BackboneProgram, attach_backbone_params, conv_bn_layer
from i2r_b200.ops import ConvLayer, Runner
from i2r_b200.packing import deconv4x4s2_phase_taps, fold_bn
from i2r_b200.position import MaskEmbedParams, build_mask_embed_program, sine_table
from i2r_b200.module_base import DevicePathModule
from i2r_b200.modules import DeconvProgram, EncoderParams

logger = logging.getLogger(__name__)


class TransPoseH(DevicePathModule):
    flavor = "interformer_pureMulti"

    def __init__(self, cfg, **kwargs):
        super().__init__()
        extra = cfg["MODEL"]["EXTRA"]
        pre = attach_backbone_params(self, extra)
        m = cfg.MODEL
        self.trans_size = list(m.TRANS_SIZE)
        d_model = m.DIM_MODEL
        w, h = m.IMAGE_SIZE
        self.use_multi_pos = bool(m.USE_MULTI_POS)
        self.position_embedding = MaskEmbedParams(self.trans_size, d_model, mode=m.MULTI_POS_EMBEDDING,
                                                  vec_dim=d_model)
        self.reduce = nn.Conv2d(pre[-1], d_model, 1, bias=False)
        if m.POS_EMBEDDING not in ("none", "learnable", "sine"):
            raise AssertionError("POS_EMBEDDING must be none/learnable/sine")
        if m.POS_EMBEDDING == "none":
            self.pos_embedding = None
        elif m.POS_EMBEDDING == "learnable":
            self.pos_embedding = nn.Parameter(torch.randn((h // 4) * (w // 4), 1, d_model))
        else:
            self.pos_embedding = nn.Parameter(sine_table(h // 4, w // 4, d_model), requires_grad=False)
        self.global_encoder = EncoderParams(d_model, m.N_HEAD, m.DIM_FEEDFORWARD, m.ENCODER_LAYERS)
        self.deconv_with_bias = bool(extra.DECONV_WITH_BIAS)
        nl, nf, nk = extra.NUM_DECONV_LAYERS, list(extra.NUM_DECONV_FILTERS), list(extra.NUM_DECONV_KERNELS)
        assert nl == len(nf), "ERROR: num_deconv_layers is different len(num_deconv_filters)"
        assert nl == len(nk), "ERROR: num_deconv_layers is different len(num_deconv_filters)"
        mods = []
        for i in range(nl):
            if nk[i] != 4:
                raise NotImplementedError("deconv kernel %d (shipped configs use 4)" % nk[i])
            mods += [nn.ConvTranspose2d(nf[i], nf[i], 4, 2, 1, 0, bias=self.deconv_with_bias),
                     nn.BatchNorm2d(nf[i], momentum=0.1), nn.ReLU(inplace=True)]
        self.deconv_layers = nn.Sequential(*mods)
        k = extra["FINAL_CONV_KERNEL"]
        self.final_layer = nn.Conv2d(d_model, cfg["MODEL"]["NUM_JOINTS"], k, 1, 1 if k == 3 else 0)
        self.pretrained_layers = extra["PRETRAINED_LAYERS"]
        self._cfg = dict(d_model=d_model, nhead=m.N_HEAD, layers=m.ENCODER_LAYERS, num_deconv=nl, final_k=k,
                         mode=m.MULTI_POS_EMBEDDING)
        self._init_device_path(Runner)

    # ------------------------------------------------------------------ weights -> device program
    def prepare(self, device=None):
        """Fold BN, pack weights and upload (called lazily by forward; call again after loading weights)."""
        device = torch.device(device) if device is not None else self.final_layer.weight.device
        sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        c = self._cfg
        if c["mode"] not in ("conv", "res") and self.use_multi_pos:
            raise NotImplementedError("MULTI_POS_EMBEDDING=%r (kernels exist for 'conv' and 'res')" % c["mode"])
        if c["final_k"] != 1:
            raise NotImplementedError("FINAL_CONV_KERNEL=3")
        prog = type("Program", (), {})()
        prog.device = device
        prog.runner = self._runner_factory(device, 1 if self.check_impl else 0)
        prog.backbone = BackboneProgram(self, sd, device)
        prog.reduce = conv_bn_layer(sd, "reduce", None, device=device)
        prog.mask_embed = build_mask_embed_program(c["mode"], sd, "position_embedding", device) if self.use_multi_pos else None
        prog.encoder = EncoderProgram(sd, "global_encoder", c["layers"], c["d_model"], c["nhead"], device)
        prog.deconvs = [DeconvProgram(sd, "deconv_layers.%d" % (3 * i), "deconv_layers.%d" % (3 * i + 1), device)
                        for i in range(c["num_deconv"])]
        prog.head = conv_bn_layer(sd, "final_layer", None, device=device)
        self._program_ready(prog)
        return self

    # ------------------------------------------------------------------ forward
    # the four stages DevicePathModule._eager (and sharded.ShardedForward) compose
    def _stage_tokens(self, p, r, x):
        feats = p.backbone.run(r, x)
        return None, None, r.conv(p.reduce, feats[-1])             # no first-stage feature / heatmaps; [S, 16, 12, d]

    def _stage_pos(self, p, r, pos_mask, hw):
        return None if p.mask_embed is None else p.mask_embed.run(r, pos_mask, hw)

    def _stage_head(self, p, r, y, feat, heat_single):
        for dc in p.deconvs:      # the reference applies the same deconv stack twice (:774-775)
            y = dc.run(r, y)
        for dc in p.deconvs:
            y = dc.run(r, y)
        return r.conv(p.head, y, out_mode="nchw32")

    def init_weights(self, pretrained="", fixed=False, print_load_info=False):
        """Training-time initialisation (reference :779-813): N(0, 0.001) convs, identity BN, then the
        ImageNet backbone checkpoint restricted to PRETRAINED_LAYERS."""
        for mod in self.modules():
            if isinstance(mod, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.normal_(mod.weight, std=0.001)
                if mod.bias is not None:
                    nn.init.constant_(mod.bias, 0)
            elif isinstance(mod, nn.BatchNorm2d):
                nn.init.constant_(mod.weight, 1)
                nn.init.constant_(mod.bias, 0)
        if os.path.isfile(pretrained):
            ckpt = torch.load(pretrained, map_location="cpu")
            own = self.state_dict()
            keep = {}
            for name, t in ckpt.items():
                if (name.split(".")[0] in self.pretrained_layers and name in own) or self.pretrained_layers[0] == "*":
                    if fixed:
                        t.requires_grad_(False)
                    keep[name] = t
                    if print_load_info:
                        print(":: {} is loaded from {}".format(name, pretrained))
            self.load_state_dict(keep, strict=False)
        elif pretrained:
            logger.error("=> please download pre-trained models first!")
            raise ValueError("{} is not exist!".format(pretrained))
        self.invalidate()


def get_pose_net(cfg, is_train, **kwargs):
    model = TransPoseH(cfg, **kwargs)
    if is_train and cfg["MODEL"]["INIT_WEIGHTS"]:
        model.init_weights(cfg["MODEL"]["PRETRAINED"], cfg["MODEL"]["BACKBONE_FIX"])
    return model
