"""Compat extras of the `lib/models` surface, each pinned on outputs of the REAL reference
(tests/golden/make_golden.py):
  * `MULTI_POS_EMBEDDING: res` -- the resnet18-stem position embedding of the box masks (position_embedding.py:14-18,
    :90-108), used by the shipped experiments/OCHuman/interformer_ochuman_tph_192_p3_b8.yaml with USE_MULTI_POS;
  * empty `MODEL.SINGLEFORMER` -- the stand-alone HRNet backbone as token source (interformer.py:143, :291-292) -- together
    with `UPSAMPLE_TYPE: upconv` (interformer.py:25-64).
CPU: parameter surface, oracle and the emulated launch sequence against the golden; GPU: the kernels."""
import json
import os

import numpy as np
import pytest
import torch

import paths  # noqa: F401
from emulator import EmuRunner
from helpers import GOLDEN, build_model, inputs_for, load_golden
from oracle import i2r_oracle

CASES = {
    "tph_ochuman_res_ragged": ("OCHuman/interformer_ochuman_tph_192_p3_b8.yaml", (),
                               "state_dict_interformer_ochuman_tph_192_p3_b8.json"),
    "hrnet_upconv_ragged": ("crowdpose/interformer_crowdpose_tph_192_p6_b4.yaml",
                            ("MODEL.SINGLEFORMER", "", "MODEL.UPSAMPLE_TYPE", "upconv", "MODEL.INIT_WEIGHTS", False),
                            "state_dict_hrnet_upconv_ragged.json"),
}


def _outputs(out):
    return out if isinstance(out, dict) else {"out": out}


def _golden(g, k):
    return g["out_" + k] if k != "out" else g["out"]


@pytest.mark.parametrize("case", sorted(CASES))
def test_surface_oracle_and_launch_sequence_match_reference(case):
    yaml_rel, opts, sd_json = CASES[case]
    cfg, model, sd = build_model(yaml_rel, opts=opts)
    with open(os.path.join(GOLDEN, sd_json)) as f:
        ref = json.load(f)
    own = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items()}
    assert own == ref
    g = load_golden(case)
    length = [int(v) for v in g["length"]]
    x, pm = inputs_for(length)
    with torch.no_grad():
        out = _outputs(i2r_oracle.forward(sd, cfg, x, pm, length))
    for k, v in out.items():
        assert float(np.abs(v.numpy() - _golden(g, k)).max()) <= 2e-5, k
    model._runner_factory = lambda device, impl: EmuRunner()
    model.prepare("cpu")
    with torch.no_grad():
        emu = _outputs(model._eager(x, pm, length))
    assert sorted(emu) == sorted(out)
    for k, v in emu.items():
        err = float(np.abs(v.numpy() - _golden(g, k)).max())
        assert err <= 1e-3, (k, err)


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_gpu_forward_matches_reference_golden(case):
    yaml_rel, opts, _ = CASES[case]
    cfg, model, sd = build_model(yaml_rel, opts=opts)
    model = model.cuda()
    g = load_golden(case)
    length = [int(v) for v in g["length"]]
    x, pm = inputs_for(length)
    out = _outputs(model(x, pm, length))
    torch.cuda.synchronize()
    for k, v in out.items():
        err = float(np.abs(v.cpu().numpy() - _golden(g, k)).max())
        assert np.isfinite(err) and err <= 1e-3, (k, err)


@pytest.mark.gpu
def test_visualize_hooks_see_token_map_and_attention_weights():
    """visualize.py:164-175: forward hooks on `model.reduce` and `model.global_encoder.layers[i].self_attn` (output[1] =
    attention weights) -- rebuilt on demand; rows sum to one, and the hooked forward returns the usual heatmaps."""
    cfg, model, sd = build_model()
    model = model.cuda()
    length = [1]
    x, pm = inputs_for(length)
    plain = model(x, pm, length).clone()
    feats, maps = [], []
    hooks = [model.reduce.register_forward_hook(lambda m, i, o: feats.append(o))]
    hooks += [layer.self_attn.register_forward_hook(lambda m, i, o: maps.append(o[1])) for layer in model.global_encoder.layers]
    out = model(x, pm, length)
    for h in hooks:
        h.remove()
    torch.cuda.synchronize()
    assert len(feats) == 1 and tuple(feats[0].shape) == (1, 96, 16, 12)
    assert len(maps) == len(model.global_encoder.layers) and all(tuple(m.shape) == (1, 192, 192) for m in maps)
    assert all(float((m.sum(-1) - 1).abs().max()) < 1e-4 for m in maps)
    assert float((out - plain).abs().max()) <= 1e-5
    with torch.no_grad():
        taps = {}
        i2r_oracle.vanilla_forward(sd, cfg, x, pm, length, taps=taps)
    if "reduce" in taps:
        assert float((feats[0].cpu() - taps["reduce"]).abs().max()) <= 2e-2
    again = model(x, pm, length)           # hooks removed: graph replay path again
    assert float((again - plain).abs().max()) <= 1e-5
