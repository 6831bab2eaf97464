"""TEST INFRASTRUCTURE: numpy restatement of the reference's heatmap post-processing.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.

Pinned against outputs of the REAL reference functions (lib/utils/transforms.py:16-30 `flip_back`,
lib/core/inference.py:20-112 `get_max_preds` / `gaussian_blur` / `taylor` / `get_final_preds` with cv2 4.13), stored in
tests/golden/postproc.npz by tests/golden/make_golden_postproc.py."""
import numpy as np


def flip_back(output_flipped, matched_parts):
    """transforms.py:16-30: reverse the width axis, swap every matched left/right joint pair."""
    out = output_flipped[:, :, :, ::-1].copy()
    for a, b in matched_parts:
        tmp = out[:, a].copy()
        out[:, a] = out[:, b]
        out[:, b] = tmp
    return out


def flip_test_merge(output, output_flipped, matched_parts):
    """function.py:158-162."""
    return (output + flip_back(output_flipped, matched_parts)) * np.float32(0.5)


def get_max_preds(hm):
    """inference.py:20-48."""
    n, k, h, w = hm.shape
    flat = hm.reshape(n, k, -1)
    idx = np.argmax(flat, 2).reshape(n, k, 1)
    maxvals = np.amax(flat, 2).reshape(n, k, 1)
    preds = np.tile(idx, (1, 1, 2)).astype(np.float32)
    preds[:, :, 0] = preds[:, :, 0] % w
    preds[:, :, 1] = np.floor(preds[:, :, 1] / w)
    preds *= np.tile(maxvals > 0.0, (1, 1, 2)).astype(np.float32)
    return preds, maxvals


def gaussian_kernel(ksize):
    """cv2.getGaussianKernel(ksize, sigma <= 0, CV_64F): fixed tables up to 7 taps, else sigma from the size."""
    small = {1: [1.0], 3: [0.25, 0.5, 0.25], 5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
             7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125]}
    if ksize in small:
        return np.asarray(small[ksize], dtype=np.float64)
    sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(-0.5 / (sigma * sigma) * x * x)
    return k / k.sum()


def gaussian_blur(hm, ksize):
    """inference.py:73-87: per map, blur the zero-padded float64 copy, store as float32, rescale to the old maximum."""
    kern = gaussian_kernel(ksize)
    r = (ksize - 1) // 2
    out = hm.copy()
    n, k, h, w = hm.shape
    for i in range(n):
        for j in range(k):
            origin_max = np.max(hm[i, j])
            dr = np.zeros((h + 2 * r, w + 2 * r))
            dr[r:h + r, r:w + r] = hm[i, j]
            rows = np.zeros_like(dr)
            for t in range(ksize):                       # separable, zero outside the padded array's interior
                sh = t - r
                src = dr[:, max(0, sh): dr.shape[1] + min(0, sh)]
                rows[:, max(0, -sh): dr.shape[1] - max(0, sh)] += kern[t] * src
            cols = np.zeros_like(dr)
            for t in range(ksize):
                sh = t - r
                src = rows[max(0, sh): dr.shape[0] + min(0, sh), :]
                cols[max(0, -sh): dr.shape[0] - max(0, sh), :] += kern[t] * src
            out[i, j] = cols[r:h + r, r:w + r]
            out[i, j] *= origin_max / np.max(out[i, j])
    return out


def taylor(hm, coord):
    """inference.py:51-70."""
    h, w = hm.shape
    px, py = int(coord[0]), int(coord[1])
    if 1 < px < w - 2 and 1 < py < h - 2:
        dx = 0.5 * (hm[py][px + 1] - hm[py][px - 1])
        dy = 0.5 * (hm[py + 1][px] - hm[py - 1][px])
        dxx = 0.25 * (hm[py][px + 2] - 2 * hm[py][px] + hm[py][px - 2])
        dxy = 0.25 * (hm[py + 1][px + 1] - hm[py - 1][px + 1] - hm[py + 1][px - 1] + hm[py - 1][px - 1])
        dyy = 0.25 * (hm[py + 2][px] - 2 * hm[py][px] + hm[py - 2][px])
        det = float(dxx) * float(dyy) - float(dxy) ** 2
        if det != 0:
            g = np.array([float(dx), float(dy)])
            hinv = np.array([[float(dyy), -float(dxy)], [-float(dxy), float(dxx)]]) / det
            coord = coord + (-hinv @ g).astype(coord.dtype)
    return coord


def transform_preds(coords, center, scale, output_size):
    """transforms.py:50-92 with rot = 0, inv = 1: a similarity about the centres (only scale[0] enters)."""
    w, h = output_size
    k = (float(scale[0]) * 200.0 - 1.0) / (w - 1.0)
    out = np.zeros(coords.shape)
    out[:, 0] = center[0] + k * (coords[:, 0] - (w - 1.0) * 0.5)
    out[:, 1] = center[1] + k * (coords[:, 1] - (h - 1.0) * 0.5)
    return out


def get_final_preds(hm, center, scale, blur_kernel, transform_back=True):
    """inference.py:90-112."""
    coords, maxvals = get_max_preds(hm)
    h, w = hm.shape[2], hm.shape[3]
    hm = gaussian_blur(hm, blur_kernel)
    hm = np.log(np.maximum(hm, 1e-10))
    for n in range(coords.shape[0]):
        for p in range(coords.shape[1]):
            coords[n, p] = taylor(hm[n][p], coords[n][p])
    preds = coords.copy()
    if transform_back:
        for i in range(coords.shape[0]):
            preds[i] = transform_preds(coords[i], center[i], scale[i], [w, h])
    return preds, maxvals
