"""BASELINE config C3 at full size on the GPU: two-stage I2R-Net with the TransPose-H first stage, 4 images x 4
persons = 16 crops, against the pinned oracle.  Tolerance 1e-3 max-abs on the fp32 heatmaps (north_star)."""
import json
import os

import pytest
import torch

import paths
from helpers import build_model, inputs_for

pytestmark = pytest.mark.gpu


def test_c3_batch16_matches_oracle():
    from oracle import i2r_oracle
    cfg, model, sd = build_model("coco/interformer_coco_tph_192_p4_b4.yaml")
    model = model.cuda()
    length = [4] * 4
    x, pm = inputs_for(length)
    out = model(x, pm, length)
    torch.cuda.synchronize()
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        ref = i2r_oracle.forward(sd, cfg, x, pm, length)
    errs = {k: float((out[k].cpu() - ref[k]).abs().max()) for k in ref}
    os.makedirs(os.path.join(paths.REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(paths.REPO, "gpurun_out", "model_report.jsonl"), "a") as f:
        f.write(json.dumps({"test": "c3_batch16", "max_abs_err": errs, "out_max": float(ref["multi"].abs().max())}) + "\n")
    assert all(v <= 1e-3 for v in errs.values()), errs
