"""Two-stage I2R-Net, `interformer_2stage` naming (reference lib/models/interformer_2stage.py:208-434; used by
experiments/coco/interformer_coco_tph_192_p4_b4.yaml).  Drop-in: same `get_pose_net(cfg, is_train)`, same
state_dict keys, `forward(x, pos_mask, length)` -> {'single', 'multi'} heatmaps (or 'multi' alone when
SINGLEFORMER_FIX / not INTER_SUPERVISION), computed by the sm_100a kernels."""
import models
from i2r_b200.two_stage import TwoStageInterFormer as InterFormer  # noqa: F401
from i2r_b200.two_stage import build


def get_pose_net(cfg, is_train, **kwargs):
    return build(cfg, is_train, "interformer_2stage", models)
