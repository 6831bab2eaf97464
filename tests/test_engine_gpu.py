"""Engine behaviour under the reference's real eval loop (lib/core/function.py:126-135): `length` changes on almost
every batch while the number of crops per batch stays the DataLoader's batch size, so graphs must be keyed on the
shape, not on the persons-per-image list (VERDICT r01 weak #10 / ADVICE medium)."""
import random

import pytest
import torch

import paths  # noqa: F401
from helpers import build_model, inputs_for
from i2r_b200.engine import HostPipeline

pytestmark = pytest.mark.gpu


def _random_partition(total, rng):
    parts, left = [], total
    while left > 0:
        n = rng.randint(1, min(left, 5))
        parts.append(n)
        left -= n
    return parts


def test_fifty_length_lists_of_equal_size_take_at_most_two_captures():
    cfg, model, _ = build_model()
    model = model.cuda()
    s = 12
    x, pm = inputs_for([s])
    rng = random.Random(0)
    lists = [[4, 4, 4]] + [_random_partition(s, rng) for _ in range(49)] + [[s], [1] * s]
    assert len({tuple(l) for l in lists}) >= 30
    model.use_cuda_graph = True
    graphed = [model(x, pm, l) for l in lists]
    torch.cuda.synchronize()
    assert model._graphs.captures <= 2, model._graphs.captures
    model.use_cuda_graph = False
    for l, g in list(zip(lists, graphed))[::5] + [(lists[-2], graphed[-2]), (lists[-1], graphed[-1])]:
        e = model(x, pm, l)
        torch.cuda.synchronize()
        # same kernels on the same data; only launch bounds differ (empty CTAs; a different key-split count moves the
        # fp16 rounding points of the attention probabilities, well inside the 1e-3 bar)
        assert float((e - g).abs().max()) <= 5e-4, (l, float((e - g).abs().max()))


def test_host_inputs_and_pipeline_match_device_inputs():
    cfg, model, _ = build_model()
    model = model.cuda()
    length = [2, 1]
    batches = [inputs_for(length, seed=s) for s in (1, 2, 3, 4, 5)]
    ref = [model(x.cuda(), pm.cuda(), length).cpu() for x, pm in batches]
    # host tensors through forward (staged upload), pageable and pinned
    for (x, pm), r in zip(batches, ref):
        assert torch.equal(model(x, pm, length).cpu(), r)
        assert torch.equal(model(x.pin_memory(), pm.pin_memory(), length).cpu(), r)
    # pipelined host-to-host loop, results consumed one step late
    pipe = HostPipeline(model, depth=2)
    got, prev = [], None
    for x, pm in batches:
        t = pipe.submit(x.pin_memory(), pm.pin_memory(), length)
        if prev is not None:
            got.append(prev.result().clone())
        prev = t
    got.append(prev.result().clone())
    for g, r in zip(got, ref):
        assert torch.equal(g, r)


def test_weight_edits_through_submodules_invalidate_the_device_program():
    """The reference loads first-stage weights through `model.singleformer.load_state_dict` (interformer.py:147-155) and
    edits parameters in place; the wrapper must not keep computing with stale packed weights."""
    cfg, model, sd = build_model("crowdpose/interformer_crowdpose_tph_192_p6_b4.yaml")
    model = model.cuda()
    length = [1, 1]
    x, pm = inputs_for(length)
    a = {k: v.clone() for k, v in model(x, pm, length).items()}
    sub = {k[len("singleformer."):]: v * 1.5 if k.endswith("final_layer.weight") else v
           for k, v in sd.items() if k.startswith("singleformer.")}
    model.singleformer.load_state_dict(sub)
    b = model(x, pm, length)
    torch.cuda.synchronize()
    assert float((b["single"] - a["single"]).abs().max()) > 1e-3        # first-stage head changed
    with torch.no_grad():
        model.final_layer.bias.add_(0.25)                                # in-place edit of a wrapper parameter
    c = model(x, pm, length)
    torch.cuda.synchronize()
    assert abs(float((c["multi"] - b["multi"]).mean()) - 0.25) < 1e-3
