"""`models.backbone.build_backbone(cfg)` (reference lib/models/backbone.py:8-25): the stand-alone HRNet wrapped so that
its parameters live under `body.*`.  Note the reference builds it with is_train=True (:11), i.e. `init_weights` runs
whenever cfg.MODEL.INIT_WEIGHTS is set -- reproduced."""
import torch.nn as nn

from models.hrnet import get_pose_net


class HRNetBackbone(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.body = get_pose_net(cfg, is_train=True)

    def forward(self, x):
        return self.body(x)


def build_backbone(cfg):
    return HRNetBackbone(cfg)
