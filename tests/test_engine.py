"""Host logic of the graph engine (no GPU): sequence plans as device DATA and the two-captures-per-shape policy."""
import numpy as np
import pytest
import torch

import paths  # noqa: F401
from emulator import EmuRunner
from helpers import build_model, inputs_for, load_golden
from i2r_b200.engine import ExactPlans, GraphedForward, SeqPlan


def test_seq_plan_offsets_are_padded_with_empty_sequences():
    plan = SeqPlan("cpu", 7, 5, 7)
    assert plan.fits([3, 1, 3]) and plan.fits([2, 1, 1, 1, 2]) and not plan.fits([1] * 7) and not plan.fits([3, 3])
    plan.update([3, 1, 3])
    assert plan.cu(192).tolist() == [0, 576, 768, 1344, 1344, 1344]
    assert plan.max_seqlen(192) == 7 * 192
    plan.update([2, 5])
    assert plan.cu(192).tolist() == [0, 384, 1344, 1344, 1344, 1344]
    assert plan.cu(432).tolist() == [0, 864, 3024, 3024, 3024, 3024]
    with pytest.raises(ValueError):
        plan.update([4, 4])


def test_capture_policy_tight_first_then_generic():
    g = GraphedForward(lambda *a, **k: None)
    skey = ((12, 3, 256, 192), (12, 1, 256, 192), 0)
    assert g._bounds_for_miss(skey, [4, 4, 4]) == (3, 4)

    class E:
        pass
    e = E()
    e.plan = SeqPlan("cpu", 12, 3, 4)
    g.entries[(skey, (3, 4))] = e
    assert g._select(skey, [4, 4, 4]) is e and g._select(skey, [3, 4, 4, 1]) is None
    assert g._select(skey, [4, 4, 2, 2]) is None and g._select(skey, [6, 6]) is None
    assert g._bounds_for_miss(skey, [6, 6]) == (12, 12)          # the generic plan: fits every list of 12 crops
    e2 = E()
    e2.plan = SeqPlan("cpu", 12, 12, 12)
    g.entries[(skey, (12, 12))] = e2
    assert g._select(skey, [6, 6]) is e2 and g._select(skey, [1] * 12) is e2
    assert g._select(skey, [4, 4, 4]) is e                        # the tight entry stays preferred


def test_exact_plans_are_bounded():
    plans = ExactPlans("cpu", capacity=4)
    for n in range(1, 10):
        plans.get([n, 1])
    assert len(plans.plans) == 4 and plans.get([9, 1]).length == [9, 1]


def test_launch_sequence_with_generic_plan_matches_exact_plan():
    """Padded offsets / loose launch bounds are data-only changes: the emulated forward is bit-identical."""
    cfg, model, _ = build_model()
    model._runner_factory = lambda device, impl: EmuRunner()
    model.prepare("cpu")
    g = load_golden("vanilla_ragged")
    length = [int(v) for v in g["length"]]
    x, pm = inputs_for(length)
    plan = SeqPlan("cpu", sum(length), sum(length), sum(length))
    plan.update(length)
    with torch.no_grad():
        exact = model._eager(x, pm, length)
        generic = model._eager(x, pm, plan)
    assert torch.equal(exact, generic)
    assert float(np.abs(generic.numpy() - g["out"]).max()) <= 1e-3
