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
