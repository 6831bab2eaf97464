"""Input pipeline on the device: person crops by affine warp + normalisation + box masks + batch
concatenation.  CPU: the numpy oracle against outputs of the REAL `JointsDataset.__getitem__` + `collater`
(tests/golden/preproc.npz, tests/golden/make_golden_preproc.py) and, where cv2 is importable, its 8-bit fixed-point
kernels against cv2 itself.  GPU: the sm_100a kernels against the oracle / golden.
Bars: the warped uint8 crops and therefore `x` are BIT-EXACT (integer arithmetic, then three IEEE float32 operations);
`pos_mask` is exact up to one grey level (1/255) on a few up-sampled pixels where cv2.resize's vectorised path rounds
differently from its own scalar formula (documented in oracle/preproc_oracle.py)."""
import numpy as np
import pytest
import torch

import paths  # noqa: F401
from helpers import load_golden
from i2r_b200.synth import synth_image, synth_people
from oracle import preproc_oracle as po

CASES = {"192": (192, 256), "288": (288, 384)}
IMAGES = [(427, 640, 3), (480, 381, 2)]


def _scene(tag):
    g = load_golden("preproc")
    images, annos = [], []
    for k, (h, w, n) in enumerate(IMAGES):
        images.append(synth_image(h, w, seed=10 * int(tag) + k))
        people = synth_people(h, w, n, seed=100 * int(tag) + k)
        assert np.array_equal(np.asarray([p["box"] for p in people]), g["box_%s_%d" % (tag, k)]), "generator drifted"
        annos.append(people)
    return g, images, annos


@pytest.mark.parametrize("tag", ["192", "288"])
def test_oracle_pipeline_matches_reference_dataset_and_collater(tag):
    g, images, annos = _scene(tag)
    size = CASES[tag]
    for k, people in enumerate(annos):
        for i, p in enumerate(people):
            t = po.get_affine_transform(p["center"], p["scale"], size)
            assert float(np.abs(t - g["trans_%s_%d" % (tag, k)][i]).max()) <= 1e-9
    x, pm, length = po.collate(images, annos, size)
    assert length == g["length_" + tag].tolist()
    assert np.array_equal(x[:, :, ::4, ::4], g["x_sub_" + tag])                       # bit-exact
    assert float(x.astype(np.float64).sum()) == float(g["x_sum_" + tag][0])
    assert float(np.abs(x).astype(np.float64).sum()) == float(g["x_sum_" + tag][1])
    d = np.abs(pm[:, :, ::4, ::4] - g["pm_sub_" + tag])
    assert float(d.max()) <= 1.0 / 255 + 1e-7 and float((d > 0).mean()) < 0.01
    assert abs(float(pm.astype(np.float64).sum()) - float(g["pm_sum_" + tag][0])) <= 64.0 / 255


def test_fixed_point_kernels_match_cv2_when_available():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for trial in range(6):
        h, w = int(rng.integers(150, 500)), int(rng.integers(150, 600))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        c = np.array([rng.uniform(0, w), rng.uniform(0, h)], dtype=np.float32)
        s = np.array([1.0, 1.25], dtype=np.float32) * np.float32(rng.uniform(0.3, 2.5))
        for size in CASES.values():
            t = po.get_affine_transform(c, s, size)
            assert np.array_equal(po.warp_affine_u8(img, t, size), cv2.warpAffine(img, t, size, flags=cv2.INTER_LINEAR))
        m = po.box_rectangle((h, w), (rng.uniform(0, w / 2), rng.uniform(0, h / 2), rng.uniform(10, w / 2), rng.uniform(10, h / 2)))
        for size in CASES.values():
            d = np.abs(po.resize_linear_u8(m, size).astype(int) - cv2.resize(m, size).astype(int))
            assert d.max() <= 1 and (d > 0).mean() < 0.01


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["192", "288"])
def test_device_pipeline_matches_oracle_and_reference_golden(tag):
    from i2r_b200.preproc import GpuCropper
    g, images, annos = _scene(tag)
    size = CASES[tag]
    x, pm, length = GpuCropper(size, "cuda:0")(images, annos)
    torch.cuda.synchronize()
    ox, opm, olen = po.collate(images, annos, size)
    assert length == olen == g["length_" + tag].tolist()
    x, pm = x.cpu().numpy(), pm.cpu().numpy()
    assert x.shape == ox.shape and pm.shape == opm.shape
    assert np.array_equal(x, ox)                                              # bit-exact against the oracle ...
    assert np.array_equal(x[:, :, ::4, ::4], g["x_sub_" + tag])              # ... and the real dataset + collater
    assert np.array_equal(pm, opm)
    d = np.abs(pm[:, :, ::4, ::4] - g["pm_sub_" + tag])
    assert float(d.max()) <= 1.0 / 255 + 1e-7 and float((d > 0).mean()) < 0.01


@pytest.mark.gpu
def test_device_pipeline_feeds_the_forward():
    """pipeline -> forward -> decode, everything on the device (N3 -> path -> N2)."""
    import types
    from core.inference import get_final_preds
    from helpers import build_model
    from i2r_b200.preproc import GpuCropper
    g, images, annos = _scene("192")
    x, pm, length = GpuCropper(CASES["192"], "cuda:0")(images, annos)
    cfg, model, _ = build_model()
    model = model.cuda()
    out = model(x, pm, length)
    centers = np.stack([p["center"] for a in annos for p in a])
    scales = np.stack([p["scale"] for a in annos for p in a])
    cfgd = types.SimpleNamespace(TEST=types.SimpleNamespace(BLUR_KERNEL=11))
    preds, maxvals = get_final_preds(cfgd, out, centers, scales)
    torch.cuda.synchronize()
    assert tuple(preds.shape) == (5, 17, 2) and bool(torch.isfinite(preds).all())
