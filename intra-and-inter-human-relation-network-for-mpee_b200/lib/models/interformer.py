"""Two-stage I2R-Net, `interformer` naming (reference lib/models/interformer.py:129-330; used by the CrowdPose /
OCHuman TransPose-H and all HRFormer yamls).  Drop-in: same `get_pose_net(cfg, is_train)`, same state_dict keys,
`forward(x, pos_mask, length)` -> {'single', 'multi'} heatmaps (or 'multi' alone when SINGLEFORMER_FIX / not
INTER_SUPERVISION), computed by the sm_100a kernels."""
import models
from i2r_b200.two_stage import TwoStageInterFormer as InterFormer  # noqa: F401
from i2r_b200.two_stage import build


def get_pose_net(cfg, is_train, **kwargs):
    return build(cfg, is_train, "interformer", models)
