"""`models.hrnet` (the stand-alone HRNet-W48-S backbone, lib/models/hrnet.py:275-487) and
`models.backbone.build_backbone` (lib/models/backbone.py): parameter surface against the reference's state_dict key
list, forward on the GPU against the committed output of the REAL reference module."""
import json
import os

import numpy as np
import pytest
import torch

import paths  # noqa: F401
from helpers import GOLDEN, load_golden
from i2r_b200.config import load_experiment
from i2r_b200.synth import synth_inputs, synth_state_dict


def _build():
    import models
    cfg = load_experiment("coco/interformer_coco_w48_pure_en6.yaml")
    model = models.hrnet.get_pose_net(cfg, is_train=False)
    model.load_state_dict(synth_state_dict(model.state_dict(), seed=0), strict=True)
    return cfg, model.eval()


def test_hrnet_state_dict_surface_matches_reference():
    cfg, model = _build()
    with open(os.path.join(GOLDEN, "state_dict_hrnet.json")) as f:
        ref = json.load(f)
    own = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items()}
    assert own == ref
    import models
    bb = models.backbone.build_backbone(cfg)
    assert sorted(bb.state_dict()) == sorted("body." + k for k in ref)
    with pytest.raises(Exception):
        model(torch.zeros(1, 3, 256, 192))          # no CPU forward


@pytest.mark.gpu
def test_hrnet_forward_matches_reference_golden():
    cfg, model = _build()
    model = model.cuda()
    x, _ = synth_inputs(1, 256, 192, seed=1)
    out = model(x)
    torch.cuda.synchronize()
    ref = load_golden("hrnet_c1")["out"]
    assert out.dtype == torch.float32 and tuple(out.shape) == ref.shape
    err = float(np.abs(out.cpu().numpy() - ref).max())
    assert err <= 5e-3 * float(np.abs(ref).max()), err      # fp16 token map (2.9 max): the heatmap bar sits after the head
