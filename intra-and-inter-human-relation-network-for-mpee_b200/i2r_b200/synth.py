"""Deterministic synthetic weights and inputs (there are no datasets or checkpoints offline).

`synth_state_dict` fills a state_dict *by key name*, independent of module construction order, so
the reference model (when generating golden vectors), the oracle and the B200 module all receive
bit-identical fp32 weights.  Scales follow PyTorch's default initialisers (conv/linear
U(+-1/sqrt(fan_in)), encoder matrices xavier-uniform) so activations stay O(1); BatchNorm /
LayerNorm statistics are made non-trivial so folding errors cannot hide.
"""
import zlib

import numpy as np
import torch


def _rng(seed, key):
    return np.random.default_rng([int(seed), zlib.crc32(key.encode("utf-8"))])


def synth_state_dict(reference_sd, seed=0):
    """Return a new state_dict with the keys/shapes/dtypes of `reference_sd` and synthetic values."""
    out = {}
    keys = list(reference_sd.keys())
    keyset = set(keys)
    for key in keys:
        ref = reference_sd[key]
        shape = tuple(ref.shape)
        leaf = key.rsplit(".", 1)[-1]
        prefix = key.rsplit(".", 1)[0] if "." in key else ""
        g = _rng(seed, key)
        if leaf == "num_batches_tracked":
            out[key] = torch.zeros(shape, dtype=ref.dtype)
            continue
        if leaf in ("pos_embedding",) or not ref.dtype.is_floating_point:
            out[key] = ref.detach().clone()  # structural constants (sine table) stay as constructed
            continue
        is_norm = (prefix + ".running_var") in keyset or ("norm" in prefix.rsplit(".", 1)[-1] and len(shape) == 1)
        if leaf == "running_var":
            val = g.uniform(0.5, 1.5, size=shape)
        elif leaf == "running_mean":
            val = g.normal(0.0, 0.1, size=shape)
        elif is_norm and leaf == "weight":
            val = g.uniform(0.5, 1.5, size=shape)
        elif is_norm and leaf == "bias":
            val = g.normal(0.0, 0.1, size=shape)
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            if "encoder" in key and len(shape) == 2:
                bound = float(np.sqrt(6.0 / (shape[0] + shape[1])))  # xavier-uniform
            else:
                bound = 1.0 / float(np.sqrt(fan_in))
            val = g.uniform(-bound, bound, size=shape)
        else:
            val = g.uniform(-0.1, 0.1, size=shape)
        out[key] = torch.from_numpy(np.asarray(val, dtype=np.float32)).to(ref.dtype)
    return out


def synth_inputs(num_crops, height=256, width=192, seed=1):
    """x ~ N(0,1) [S,3,H,W] and pos_mask ~ U(0,1) [S,1,H,W], fp32."""
    g = np.random.default_rng(int(seed))
    x = g.standard_normal(size=(num_crops, 3, height, width), dtype=np.float32)
    pm = g.random(size=(num_crops, 1, height, width), dtype=np.float32)
    return torch.from_numpy(x), torch.from_numpy(pm)


def synth_heatmaps(n, k, h, w, seed=0):
    """Heatmap-like fp32 maps [n,k,h,w], bit-reproducible on every platform (integers from PCG64 and IEEE add / mul /
    div only, no transcendental functions): one to three rational bumps amp / (1 + d^2 / s^2) plus small noise; map
    (0,0) is entirely negative, (0,1) has its maximum in a corner, (0,2) one pixel outside the Taylor interior."""
    g = np.random.default_rng(int(seed))
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    hm = np.zeros((n, k, h, w), dtype=np.float32)
    for i in range(n):
        for j in range(k):
            m = (g.integers(-1000, 1001, size=(h, w)).astype(np.float64)) / 100000.0
            for _ in range(int(g.integers(1, 4))):
                cy = float(g.integers(-200, 100 * h + 200)) / 100.0
                cx = float(g.integers(-200, 100 * w + 200)) / 100.0
                amp = float(g.integers(20, 101)) / 100.0
                s2 = (float(g.integers(100, 301)) / 100.0) ** 2
                m = m + amp / (1.0 + ((ys - cy) ** 2 + (xs - cx) ** 2) / (2.0 * s2))
            hm[i, j] = m.astype(np.float32)
    hm[0, 0] = -np.abs(hm[0, 0]) - np.float32(0.01)
    hm[0, 1, 0, 0] = 5.0
    hm[0, 2, h - 2, w - 3] = 5.0
    return hm


def synth_image(h, w, seed=0):
    """Bit-reproducible RGB uint8 image [h,w,3]: smooth integer gradients, blocks and integer noise (PCG64 integers and
    integer arithmetic only)."""
    g = np.random.default_rng(int(seed))
    ys, xs = np.mgrid[0:h, 0:w]
    img = np.zeros((h, w, 3), dtype=np.int64)
    for c in range(3):
        a, b, d = (int(v) for v in g.integers(1, 7, size=3))
        img[:, :, c] = (a * xs + b * ys + d * ((xs // 16 + ys // 16) % 2) * 40) % 256
    for _ in range(12):
        y0, x0 = int(g.integers(0, h - 20)), int(g.integers(0, w - 20))
        hh, ww = int(g.integers(10, h // 3)), int(g.integers(10, w // 3))
        img[y0:y0 + hh, x0:x0 + ww] = g.integers(0, 256, size=3)
    img = (img + g.integers(-12, 13, size=img.shape)).clip(0, 255)
    return img.astype(np.uint8)


def synth_people(h, w, n, seed=0):
    """n person boxes inside an h x w image with the center / scale the reference derives from a box
    (lib/dataset/coco.py `_box2cs`: aspect ratio 192/256, pixel_std 200, 1.25 enlargement)."""
    g = np.random.default_rng(int(seed))
    people = []
    for _ in range(n):
        bw, bh = float(g.integers(40, w // 2)), float(g.integers(60, h // 2))
        x, y = float(g.integers(0, w - int(bw))) + 0.25 * float(g.integers(0, 4)), float(g.integers(0, h - int(bh)))
        center = np.array([x + bw * 0.5, y + bh * 0.5], dtype=np.float32)
        aspect = 192.0 / 256.0
        if bw > aspect * bh:
            bh2, bw2 = bw / aspect, bw
        else:
            bh2, bw2 = bh, bh * aspect
        scale = np.array([bw2 / 200.0, bh2 / 200.0], dtype=np.float32) * 1.25
        people.append(dict(box=(x, y, bw, bh), center=center, scale=scale))
    return people
