"""Shared test helpers: build the drop-in module + synthetic weights for a BASELINE config."""
import os

import numpy as np
import torch

import paths  # noqa: F401
from i2r_b200.config import load_experiment
from i2r_b200.synth import synth_inputs, synth_state_dict

GOLDEN = os.path.join(paths.REPO, "tests", "golden")
VANILLA_YAML = "coco/interformer_coco_w48_pure_en6.yaml"


def build_model(yaml_rel=VANILLA_YAML, seed=0, opts=()):
    """Drop-in module with synthetic weights (same weights the golden files were generated with)."""
    import models  # the repo's lib/models
    cfg = load_experiment(yaml_rel, opts)
    torch.manual_seed(0)
    model = eval("models." + cfg.MODEL.NAME + ".get_pose_net")(cfg, is_train=False)
    sd = synth_state_dict(model.state_dict(), seed=seed)
    model.load_state_dict(sd, strict=True)
    return cfg, model.eval(), sd


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def inputs_for(length, h=256, w=192, seed=1):
    return synth_inputs(sum(length), h, w, seed=seed)
