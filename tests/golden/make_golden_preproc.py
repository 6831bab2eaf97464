"""Generate tests/golden/preproc.npz from the REAL reference input pipeline (build container only; needs
/root/reference, cv2, torchvision):  python tests/golden/make_golden_preproc.py

Two synthetic RGB images (i2r_b200.synth.synth_image, bit-reproducible) are written as PNG, a JointsDataset subclass with
a hand-made db (three and two persons) runs the reference's own `__getitem__` (is_train False) with the transform of
tools/test.py:126-134, and the reference `collater(0)` concatenates the batch.  Stored: the annotations, the affine
matrices, and the outputs subsampled (every 4th row / column) plus full-resolution checksums."""
import os
import sys
import types

import cv2
import numpy as np
import torch
import torchvision.transforms as transforms

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("I2R_REF", "/root/reference")
sys.path.insert(0, os.path.join(REPO, "oracle", "ref_shims"))
sys.path.insert(0, os.path.join(REF, "lib"))
sys.path.insert(0, os.path.join(REPO, "intra-and-inter-human-relation-network-for-mpee_b200"))
import importlib.util  # noqa: E402


def _load(name, rel):
    """lib/dataset/__init__.py imports every dataset (pycocotools, json_tricks ... are absent): load the two files the
    input pipeline consists of directly."""
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, "lib", rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


JointsDataset = _load("ref_JointsDataset", "dataset/JointsDataset.py").JointsDataset  # (reference)
collater = _load("ref_collater", "dataset/collater.py").collater  # (reference)
from utils.transforms import get_affine_transform  # noqa: E402  (reference)
from i2r_b200.synth import synth_image, synth_people  # noqa: E402


def cfg_for(image_size, heatmap_size):
    ns = types.SimpleNamespace
    return ns(OUTPUT_DIR="", DATASET=ns(DATA_FORMAT="jpg", SCALE_FACTOR=0.3, ROT_FACTOR=40, FLIP=False,
                                         NUM_JOINTS_HALF_BODY=8, PROB_HALF_BODY=0.0, COLOR_RGB=True),
              MODEL=ns(TARGET_TYPE="gaussian", IMAGE_SIZE=list(image_size), HEATMAP_SIZE=list(heatmap_size), SIGMA=2),
              LOSS=ns(USE_DIFFERENT_JOINTS_WEIGHT=False))


class TinyDataset(JointsDataset):
    def __init__(self, cfg, db, transform):
        super().__init__(cfg, "", "val", False, transform)
        self.num_joints = 17
        self.db = db


def main():
    out = {}
    tmp = "/tmp/i2r_preproc_golden"
    os.makedirs(tmp, exist_ok=True)
    normalize = transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    transform = transforms.Compose([transforms.ToTensor(), normalize])
    for tag, (size, hsize) in {"192": ((192, 256), (48, 64)), "288": ((288, 384), (72, 96))}.items():
        db = []
        for k, (h, w, n) in enumerate([(427, 640, 3), (480, 381, 2)]):
            img = synth_image(h, w, seed=10 * int(tag) + k)
            path = os.path.join(tmp, "img_%s_%d.png" % (tag, k))
            cv2.imwrite(path, img[:, :, ::-1])              # the dataset reads BGR and converts to RGB
            people = synth_people(h, w, n, seed=100 * int(tag) + k)
            annos = [dict(joints_3d=np.zeros((17, 3), np.float32), joints_3d_vis=np.zeros((17, 3), np.float32),
                          center=p["center"].copy(), scale=p["scale"].copy(), box=list(p["box"]), score=1, imgnum=i) for i, p in enumerate(people)]
            db.append(dict(image=path, annos=annos))
            out["box_%s_%d" % (tag, k)] = np.asarray([p["box"] for p in people], dtype=np.float64)
            out["center_%s_%d" % (tag, k)] = np.stack([p["center"] for p in people])
            out["scale_%s_%d" % (tag, k)] = np.stack([p["scale"] for p in people])
            out["trans_%s_%d" % (tag, k)] = np.stack([get_affine_transform(p["center"], p["scale"], 0, np.array(size))
                                                       for p in people])
        ds = TinyDataset(cfg_for(size, hsize), db, transform)
        batch = [ds[i] for i in range(len(db))]
        inp, pos_mask, _, _, meta = collater(0)(batch)
        out["length_" + tag] = meta["length"].numpy()
        x, pm = inp.numpy(), pos_mask.numpy()
        out["x_sub_" + tag], out["pm_sub_" + tag] = x[:, :, ::4, ::4].copy(), pm[:, :, ::4, ::4].copy()
        out["x_sum_" + tag] = np.asarray([float(x.astype(np.float64).sum()), float(np.abs(x).astype(np.float64).sum())])
        out["pm_sum_" + tag] = np.asarray([float(pm.astype(np.float64).sum())])
        # the warped uint8 crops themselves (before ToTensor / Normalize), for the bit-exactness check of the warp
        crops = np.rint((x * np.array([0.229, 0.224, 0.225], np.float32)[None, :, None, None] +
                         np.array([0.485, 0.456, 0.406], np.float32)[None, :, None, None]) * 255).astype(np.int64)
        out["crop_hist_" + tag] = np.bincount(crops.reshape(-1).clip(0, 255), minlength=256)
    np.savez_compressed(os.path.join(HERE, "preproc.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
