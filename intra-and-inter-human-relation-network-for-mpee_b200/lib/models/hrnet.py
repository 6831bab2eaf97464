"""Stand-alone HRNet-W48-S backbone ("hrnet"): stem + layer1 + two multi-resolution stages + `reduce` 1x1 on the
lowest-resolution branch -> token map [S, d_model, H/16, W/16].

Drop-in for the reference module of the same name (lib/models/hrnet.py:275-487; reached through
lib/models/backbone.py:11 when cfg.MODEL.SINGLEFORMER is empty, lib/models/interformer.py:143): same factory signature
`get_pose_net(cfg, is_train, **kw)`, same state_dict keys/shapes (including the `final_layer` the reference constructs
but never applies, :317-323), `forward(x)` computed by the sm_100a kernels of libi2r_sm100.so; no CPU forward.
"""
import logging
import os

import torch
import torch.nn as nn

from i2r_b200 import capi
from i2r_b200.hrnet_w48 import BackboneProgram, attach_backbone_params, conv_bn_layer
from i2r_b200.ops import Runner

logger = logging.getLogger(__name__)


class HRNet(nn.Module):
    def __init__(self, cfg, **kwargs):
        super().__init__()
        extra = cfg["MODEL"]["EXTRA"]
        pre = attach_backbone_params(self, extra)
        d_model = cfg.MODEL.DIM_MODEL
        self.reduce = nn.Conv2d(pre[-1], d_model, 1, bias=False)
        k = extra["FINAL_CONV_KERNEL"]
        self.final_layer = nn.Conv2d(d_model, cfg["MODEL"]["NUM_JOINTS"], k, 1, 1 if k == 3 else 0)
        self.pretrained_layers = extra["PRETRAINED_LAYERS"]
        self.precision = "fp16"
        self._program = None
        self._runner = None

    def build_program(self, device):
        """(BackboneProgram, reduce layer) on `device`: what forward -- or a wrapper that owns this module -- runs."""
        sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        prog = type("Program", (), {})()
        prog.device = torch.device(device)
        prog.backbone = BackboneProgram(self, sd, prog.device)
        prog.reduce = conv_bn_layer(sd, "reduce", None, device=prog.device)
        return prog

    def load_state_dict(self, *a, **kw):
        out = super().load_state_dict(*a, **kw)
        self._program = None
        return out

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._program = None
        return out

    def forward(self, x):
        dev = self.reduce.weight.device
        if dev.type != "cuda":
            raise capi.I2RError("hrnet forward runs on a CUDA (sm_100a) device only; move the module with .cuda() -- "
                                "there is no CPU fallback")
        if self._program is None or self._program.device != dev:
            self._program = self.build_program(dev)
            self._runner = Runner(dev, 0)
        x = x.to(dev, dtype=torch.float32, non_blocking=True).contiguous()
        with torch.no_grad(), torch.cuda.device(dev):
            feats = self._program.backbone.run(self._runner, x)
            tok = self._runner.conv(self._program.reduce, feats[-1])       # fp16 NHWC [S, h, w, d]
            return tok.permute(0, 3, 1, 2).float()                          # the reference returns NCHW fp32

    def init_weights(self, pretrained="", print_load_info=False):
        """Reference :448-479: N(0, 0.001) convs, identity BN, then the PRETRAINED_LAYERS of a checkpoint."""
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
                    keep[name] = t
                    if print_load_info:
                        print(":: {} is loaded from {}".format(name, pretrained))
            self.load_state_dict(keep, strict=False)
        elif pretrained:
            logger.error("=> please download pre-trained models first!")
            raise ValueError("{} is not exist!".format(pretrained))
        self._program = None


def get_pose_net(cfg, is_train, **kwargs):
    model = HRNet(cfg, **kwargs)
    if is_train and cfg["MODEL"]["INIT_WEIGHTS"]:
        model.init_weights(cfg["MODEL"]["PRETRAINED"])
    return model
