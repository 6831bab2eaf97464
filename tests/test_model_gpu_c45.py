"""BASELINE configs C4 / C5 at their real shapes on the GPU: the HRFormer-B
two-stage I2R-Net through `models.interformer.get_pose_net` / forward,

  * C4 per rank: one image of 8 persons at 256x192 (inter-human sequence of 8 * 192 = 1536 tokens, d_model 78 -> 80),
  * C4 full batch: 8 images x 8 persons = 64 crops on one GPU,
  * C5 per rank: one image of 12 persons at 384x288 (12 * 432 = 5184 tokens; `attention_tc_kernel<80, split>` and the
    non-tail encoder branch inside the model),
  * a three-image ragged batch,

against (a) committed outputs of the REAL reference (tests/golden/make_golden.py, subsampled heatmaps) and (b) the
pinned oracle on the same inputs.  Reference semantics: lib/models/interformer.py:282-323, lib/models/hrformer.py:1138-1240.
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
Y192 = "coco/interformer_coco_hrt_192_p2_b12.yaml"
Y288 = "coco/interformer_coco_hrt_288_p2_b4.yaml"


def _report(**kw):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(json.dumps(kw) + "\n")


def _oracle(sd, cfg, x, pm, length):
    from oracle import i2r_oracle
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        return i2r_oracle.forward(sd, cfg, x, pm, length)


GOLDEN_CASES = [(Y192, "hrt_c4_rank", 256, 192), (Y288, "hrt_c5_rank", 384, 288), (Y192, "hrt_ragged3", 256, 192)]


@pytest.mark.parametrize("yaml_rel,case,h,w", GOLDEN_CASES, ids=[c[1] for c in GOLDEN_CASES])
def test_hrformer_baseline_shape_matches_reference_golden_and_oracle(yaml_rel, case, h, w):
    cfg, model, sd = build_model(yaml_rel)
    model = model.cuda()
    g = load_golden(case)
    length = [int(v) for v in g["length"]]
    sub = int(g["subsample"])
    x, pm = synth_inputs(sum(length), h, w, seed=1)
    out = model(x, pm, length)
    torch.cuda.synchronize()
    got = {k: out[k].cpu() for k in ("single", "multi")}
    err_golden = {k: float(np.abs(got[k].numpy()[:, :, ::sub, ::sub] - g["out_" + k]).max()) for k in got}
    ref = _oracle(sd, cfg, x, pm, length)
    err_oracle = {k: float((got[k] - ref[k]).abs().max()) for k in got}
    _report(test="hrt_baseline_shape", case=case, length=length, err_vs_reference_golden=err_golden,
            err_vs_oracle=err_oracle, out_absmax=float(ref["multi"].abs().max()))
    assert all(np.isfinite(v) and v <= TOL for v in err_golden.values()), err_golden
    assert all(np.isfinite(v) and v <= TOL for v in err_oracle.values()), err_oracle


def test_c4_full_batch_64_crops_matches_oracle():
    """8 images x 8 persons on one GPU (BASELINE C4's whole batch)."""
    cfg, model, sd = build_model(Y192)
    model = model.cuda()
    length = [8] * 8
    x, pm = synth_inputs(sum(length), 256, 192, seed=1)
    out = model(x, pm, length)
    torch.cuda.synchronize()
    ref = _oracle(sd, cfg, x, pm, length)
    errs = {k: float((out[k].cpu() - ref[k]).abs().max()) for k in ref}
    _report(test="c4_batch64", max_abs_err=errs, out_absmax=float(ref["multi"].abs().max()))
    assert all(np.isfinite(v) and v <= TOL for v in errs.values()), errs
    # the same crops as 8 single-image calls give the same heatmaps (images are independent)
    one = model(x[:8], pm[:8], [8])
    torch.cuda.synchronize()
    assert float((one["multi"] - out["multi"][:8]).abs().max()) <= 1e-4
