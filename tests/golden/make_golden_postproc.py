"""Generate tests/golden/postproc.npz from the REAL reference post-processing functions (run in the build container
only; needs /root/reference and cv2):  python tests/golden/make_golden_postproc.py

Heatmaps come from i2r_b200.synth.synth_heatmaps (bit-reproducible everywhere, so only a checksum is stored): bumps plus
noise, one map entirely <= 0, maxima in a corner and next to the border.  Stored: `flip_back` as the joint / column
permutation it applies plus a position-weighted checksum, and `get_final_preds` (blur kernels 3 and 11, both heatmap
sizes) as computed by lib/utils/transforms.py and lib/core/inference.py of the reference."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("I2R_REF", "/root/reference")
sys.path.insert(0, os.path.join(REF, "lib"))
from utils.transforms import flip_back  # noqa: E402  (reference)
from core.inference import get_final_preds  # noqa: E402  (reference)

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "intra-and-inter-human-relation-network-for-mpee_b200"))
from i2r_b200.synth import synth_heatmaps  # noqa: E402  (bit-reproducible generator shared with the tests)

FLIP_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]      # lib/dataset/coco.py:81-82


def main():
    rng = np.random.default_rng(0)
    out = {"flip_pairs": np.asarray(FLIP_PAIRS, dtype=np.int64)}
    for tag, (h, w) in {"192": (64, 48), "288": (96, 72)}.items():
        hm = synth_heatmaps(3, 17, h, w, seed=int(tag))
        center = rng.uniform(50, 600, (3, 2)).astype(np.float32)
        scale = rng.uniform(0.5, 3.0, (3, 2)).astype(np.float32)
        out["center_" + tag], out["scale_" + tag] = center, scale
        out["hm_checksum_" + tag] = np.asarray([float(hm.astype(np.float64).sum()), float(np.abs(hm).max())])
        fb = flip_back(hm.copy(), FLIP_PAIRS).copy()
        # flip_back is a pure permutation: joint / column index of every output element, as applied to index grids
        jidx = np.broadcast_to(np.arange(17, dtype=np.float32).reshape(1, 17, 1, 1), hm.shape).copy()
        xidx = np.broadcast_to(np.arange(w, dtype=np.float32).reshape(1, 1, 1, w), hm.shape).copy()
        out["flip_back_checksum_" + tag] = np.asarray([float((fb.astype(np.float64) * (1 + jidx) * (1 + xidx)).sum())])
        # (the reference's flip_back permutes its argument in place, so the index grids are used last)
        out["flip_back_joint_" + tag] = flip_back(jidx.copy(), FLIP_PAIRS)[0, :, 0, 0].copy()
        out["flip_back_col_" + tag] = flip_back(xidx.copy(), FLIP_PAIRS)[0, 0, 0, :].copy()
        for ks in (3, 11):
            cfg = types.SimpleNamespace(TEST=types.SimpleNamespace(BLUR_KERNEL=ks))
            preds, maxvals = get_final_preds(cfg, hm.copy(), center, scale)
            out["preds_%s_k%d" % (tag, ks)], out["maxvals_%s_k%d" % (tag, ks)] = preds, maxvals
            coords, _ = get_final_preds(cfg, hm.copy(), center, scale, transform_back=False)
            out["coords_%s_k%d" % (tag, ks)] = coords
    np.savez_compressed(os.path.join(HERE, "postproc.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
