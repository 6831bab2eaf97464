"""GPU parity of the HRFormer-B two-stage I2R-Net (configs C4 / C5 families) through
`models.interformer.get_pose_net` / forward, against the committed outputs of the REAL reference.
Tolerance: 1e-3 max-abs on fp32 heatmaps (north_star)."""
import json
import os

import numpy as np
import pytest
import torch

import paths
from helpers import build_model, load_golden
from i2r_b200.synth import synth_inputs

pytestmark = pytest.mark.gpu
TOL = 1e-3
REPORT = os.path.join(paths.REPO, "gpurun_out", "model_report.jsonl")

CASES = [
    ("coco/interformer_coco_hrt_192_p2_b12.yaml", "hrt2stage_ragged", 256, 192),
    ("coco/interformer_coco_hrt_288_p2_b4.yaml", "hrt288_c1", 384, 288),
]


@pytest.mark.parametrize("yaml_rel,case,h,w", CASES, ids=[c[1] for c in CASES])
def test_hrformer_two_stage_matches_reference_golden(yaml_rel, case, h, w):
    cfg, model, _ = build_model(yaml_rel)
    model = model.cuda()
    model.use_cuda_graph = False
    g = load_golden(case)
    length = [int(v) for v in g["length"]]
    x, pm = synth_inputs(sum(length), h, w, seed=1)
    out = model(x, pm, length)
    torch.cuda.synchronize()
    errs = {k: float(np.abs(out[k].cpu().numpy() - g["out_" + k]).max()) for k in ("single", "multi")}
    with open(REPORT, "a") as f:
        f.write(json.dumps({"test": "hrt_two_stage_golden", "case": case, "max_abs_err": errs,
                            "launches": model._program.runner.launches}) + "\n")
    assert all(np.isfinite(v) and v <= TOL for v in errs.values()), errs
    # graph replay equals eager
    model.use_cuda_graph = True
    g1 = model(x, pm, length)
    g2 = model(x, pm, length)
    torch.cuda.synchronize()
    for k in ("single", "multi"):
        assert torch.equal(g1[k], g2[k]) and torch.equal(g1[k], out[k])
