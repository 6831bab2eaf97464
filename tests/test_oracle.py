"""CPU tests: the oracle restatement against outputs of the real reference (tests/golden), and the
drop-in module's parameter surface against the reference's state_dict key/shape list."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, build_model, inputs_for, load_golden
from oracle import i2r_oracle


@pytest.fixture(scope="module")
def vanilla():
    return build_model()


def test_state_dict_surface_matches_reference(vanilla):
    _, model, _ = vanilla
    with open(os.path.join(GOLDEN, "state_dict_interformer_pureMulti.json")) as f:
        ref = json.load(f)
    own = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items()}
    assert sorted(own) == sorted(ref)
    assert own == ref
    assert len(own) == 1054  # SURVEY.md 8b [measured]


@pytest.mark.parametrize("case", ["vanilla_c1", "vanilla_ragged"])
def test_oracle_matches_reference_outputs(vanilla, case):
    cfg, _, sd = vanilla
    g = load_golden(case)
    length = [int(v) for v in g["length"]]
    x, pm = inputs_for(length)
    taps = {}
    with torch.no_grad():
        out = i2r_oracle.vanilla_forward(sd, cfg, x, pm, length, taps=taps)
    assert out.shape == g["out"].shape
    err = float(np.abs(out.numpy() - g["out"]).max())
    assert err <= 2e-5, err
    # intermediate taps: reduce output of the padded batch == reference hook output on valid crops
    assert float(np.abs(taps["reduce"].numpy() - g["tap_reduce"]).max()) <= 2e-5


def test_unpad_matches_reference_helper():
    from utils.utils import get_valid_output
    t = torch.arange(4 * 3 * 2, dtype=torch.float32).reshape(4, 3, 2)   # bs=2, N=2
    out = get_valid_output(t, [2, 1])
    assert torch.equal(out, torch.cat([t[0:2], t[2:3]], dim=0))
    assert torch.equal(out, i2r_oracle.unpad_persons(t, [2, 1]))


TWO_STAGE = [
    ("coco/interformer_coco_tph_192_p4_b4.yaml", "tph2stage_ragged", 1104),
    ("crowdpose/interformer_crowdpose_tph_192_p6_b4.yaml", "tph_crowdpose_ragged", 1062),
]


@pytest.mark.parametrize("yaml_rel,case,nkeys", TWO_STAGE, ids=[c[1] for c in TWO_STAGE])
def test_two_stage_oracle_and_surface_match_reference(yaml_rel, case, nkeys):
    """TransPose-H first stage + inter-human stage (models.interformer / models.interformer_2stage): the oracle
    restatement against outputs of the real reference, and the drop-in module's state_dict against the reference's."""
    cfg, model, sd = build_model(yaml_rel)
    with open(os.path.join(GOLDEN, "state_dict_%s.json" % os.path.basename(yaml_rel)[:-5])) as f:
        ref = json.load(f)
    own = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items()}
    assert own == ref and len(own) == nkeys
    g = load_golden(case)
    length = [int(v) for v in g["length"]]
    x, pm = inputs_for(length)
    with torch.no_grad():
        out = i2r_oracle.forward(sd, cfg, x, pm, length)
    assert sorted(out) == ["multi", "single"]
    for k in out:
        err = float(np.abs(out[k].numpy() - g["out_" + k]).max())
        assert err <= 2e-5, (k, err)


HRT = [
    ("coco/interformer_coco_hrt_192_p2_b12.yaml", "hrt2stage_ragged", 256, 192),
    ("coco/interformer_coco_hrt_288_p2_b4.yaml", "hrt288_c1", 384, 288),
]


@pytest.mark.parametrize("yaml_rel,case,h,w", HRT, ids=[c[1] for c in HRT])
def test_hrformer_oracle_matches_reference(yaml_rel, case, h, w):
    """HRFormer-B first stage (7x7 window attention incl. zero-padded tokens and WITHOUT the relative
    position bias, MlpDWBN, bilinear fuse) + the inter-human stage at d = 78: the oracle restatement against outputs
    of the REAL reference.  The weights are rebuilt from the committed key/shape list of the reference's state_dict
    (synthetic values are keyed by parameter name, so they equal the ones the golden generator loaded)."""
    from i2r_b200.config import load_experiment
    from i2r_b200.synth import synth_inputs, synth_state_dict
    cfg = load_experiment(yaml_rel)
    with open(os.path.join(GOLDEN, "state_dict_%s.json" % os.path.basename(yaml_rel)[:-5])) as f:
        ref = json.load(f)
    assert len([k for k in ref if k.startswith("singleformer.backbone.")]) == 2074      # SURVEY.md 8b [measured]
    shapes = {k: torch.zeros(v[0], dtype=getattr(torch, v[1])) for k, v in ref.items()}
    sd = synth_state_dict(shapes, seed=0)
    g = load_golden(case)
    length = [int(v) for v in g["length"]]
    x, pm = synth_inputs(sum(length), h, w, seed=1)
    with torch.no_grad():
        out = i2r_oracle.forward(sd, cfg, x, pm, length)
    for k in ("single", "multi"):
        assert out[k].shape == g["out_" + k].shape
        err = float(np.abs(out[k].numpy() - g["out_" + k]).max())
        assert err <= 5e-5, (k, err)


@pytest.mark.parametrize("yaml_rel", [h[0] for h in HRT], ids=[h[1] for h in HRT])
def test_hrformer_two_stage_surface_matches_reference(yaml_rel):
    """`models.interformer.get_pose_net` with SINGLEFORMER = hrformer: the drop-in module's state_dict (keys, shapes,
    dtypes) equals the real reference's, so its checkpoints load with strict=True."""
    cfg, model, sd = build_model(yaml_rel)
    with open(os.path.join(GOLDEN, "state_dict_%s.json" % os.path.basename(yaml_rel)[:-5])) as f:
        ref = json.load(f)
    own = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items()}
    assert sorted(own) == sorted(ref)
    assert own == ref
