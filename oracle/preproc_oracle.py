"""TEST INFRASTRUCTURE: numpy restatement of the reference's test-time input pipeline -- person
crop by affine warp, normalisation, per-person box mask, batch concatenation.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline may import this.

Follows lib/dataset/JointsDataset.py:166-201 (`get_position`, `rotate_bound`), :296-331 (`__getitem__`, is_train False),
lib/utils/transforms.py:58-92 (`get_affine_transform`), lib/dataset/collater.py:14-26,:175-181, tools/test.py:126-134
(ToTensor + Normalize), and restates the OpenCV 8-bit fixed-point kernels those lines call (cv2 is a third-party
dependency, requirements.txt: opencv-python, unpinned; container: 4.13):
  * cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0): inverse map in 1/1024 fixed point rounded to 1/32 pixel, bilinear
    weights as 15-bit integers, (sum + 2^14) >> 15   -- bit-exact against cv2 4.13 on random images (test_preproc.py);
  * cv2.getAffineTransform: the 6x6 linear system of the three point pairs, solved in double;
  * cv2.rectangle(thickness -1): inclusive corner box, clipped;
  * cv2.resize(INTER_LINEAR) on uint8: 11-bit coefficients, horizontal pass in int, vertical pass
    ((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2   -- within 1 grey level of cv2 4.13 (its IPP / SIMD path rounds a
    few up-sampled pixels differently), i.e. 1/255 on the mask.
Pinned against outputs of the REAL `JointsDataset.__getitem__` + `collater` (tests/golden/preproc.npz)."""
import numpy as np

MEAN = np.asarray([0.485, 0.456, 0.406], dtype=np.float32)      # tools/test.py:126-128
STD = np.asarray([0.229, 0.224, 0.225], dtype=np.float32)


def get_affine_transform(center, scale, output_size):
    """transforms.py:58-92 with rot = 0, shift = 0, inv = 0 (float32 point coordinates, double solve)."""
    scale_tmp = np.asarray(scale, dtype=np.float32) * np.float32(200.0)
    src_w = scale_tmp[0]
    dst_w, dst_h = output_size[0], output_size[1]
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = np.asarray(center, dtype=np.float32)
    src[1, :] = np.asarray(center, dtype=np.float32) + np.asarray([0, (src_w - 1) * -0.5])
    dst[0, :] = [(dst_w - 1) * 0.5, (dst_h - 1) * 0.5]
    dst[1, :] = np.array([(dst_w - 1) * 0.5, (dst_h - 1) * 0.5]) + np.array([0, (dst_w - 1) * -0.5], np.float32)
    for p in (src, dst):
        d = p[0] - p[1]
        p[2] = p[1] + np.array([-d[1], d[0]], dtype=np.float32)
    a = np.zeros((6, 6))
    b = np.zeros(6)
    for i in range(3):
        a[2 * i, 0:3] = [src[i, 0], src[i, 1], 1.0]
        a[2 * i + 1, 3:6] = [src[i, 0], src[i, 1], 1.0]
        b[2 * i], b[2 * i + 1] = dst[i, 0], dst[i, 1]
    return np.linalg.solve(a, b).reshape(2, 3)


def invert_affine(m):
    """cv2.warpAffine's own inversion of the forward matrix (imgwarp.cpp)."""
    m = np.asarray(m, dtype=np.float64).reshape(6).copy()
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0] = a11
    m[1] *= -d
    m[3] *= -d
    m[4] = a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m


def warp_affine_u8(src, m, dsize):
    """cv2.warpAffine(src, m, dsize, flags=cv2.INTER_LINEAR) for uint8 images, BORDER_CONSTANT 0."""
    w, h = int(dsize[0]), int(dsize[1])
    mi = invert_affine(m)
    xs, ys = np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64)
    adelta = np.rint(mi[0] * xs * 1024).astype(np.int64)
    bdelta = np.rint(mi[3] * xs * 1024).astype(np.int64)
    x0 = np.rint((mi[1] * ys + mi[2]) * 1024).astype(np.int64) + 16
    y0 = np.rint((mi[4] * ys + mi[5]) * 1024).astype(np.int64) + 16
    xq = (x0[:, None] + adelta[None, :]) >> 5
    yq = (y0[:, None] + bdelta[None, :]) >> 5
    sx, sy = np.clip(xq >> 5, -32768, 32767), np.clip(yq >> 5, -32768, 32767)
    fx, fy = xq & 31, yq & 31
    sh, sw = src.shape[:2]
    s3 = src.reshape(sh, sw, -1).astype(np.int64)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < sh) & (xx >= 0) & (xx < sw)
        return s3[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)] * ok[..., None]
    w00, w01 = (32 - fx) * (32 - fy) * 32, fx * (32 - fy) * 32
    w10, w11 = (32 - fx) * fy * 32, fx * fy * 32
    acc = (tap(sy, sx) * w00[..., None] + tap(sy, sx + 1) * w01[..., None] + tap(sy + 1, sx) * w10[..., None] +
           tap(sy + 1, sx + 1) * w11[..., None])
    return ((acc + (1 << 14)) >> 15).astype(np.uint8).reshape((h, w) + src.shape[2:])


def resize_coeffs(dn, sn):
    """Source offsets and 11-bit coefficient pairs of cv2.resize(INTER_LINEAR) along one axis."""
    scale = 1.0 / (float(dn) / sn)
    ofs = np.zeros(dn, dtype=np.int64)
    co = np.zeros((dn, 2), dtype=np.int64)
    for d in range(dn):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if s < 0:
            s, f = 0, np.float32(0)
        if s >= sn - 1:
            s, f = sn - 1, np.float32(0)
        ofs[d] = s
        co[d, 0] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))))
        co[d, 1] = int(np.rint(np.float32(f * np.float32(2048))))
    return ofs, co


def resize_linear_u8(src, dsize):
    """cv2.resize(src, dsize) (INTER_LINEAR) for a single-channel uint8 image."""
    dw, dh = int(dsize[0]), int(dsize[1])
    sh, sw = src.shape
    xo, xa = resize_coeffs(dw, sw)
    yo, ya = resize_coeffs(dh, sh)
    s = src.astype(np.int64)
    rows = s[:, xo] * xa[:, 0][None, :] + s[:, np.minimum(xo + 1, sw - 1)] * xa[:, 1][None, :]
    r0, r1 = rows[yo], rows[np.minimum(yo + 1, sh - 1)]
    out = (((ya[:, 0][:, None] * (r0 >> 4)) >> 16) + ((ya[:, 1][:, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def box_rectangle(shape, box):
    """JointsDataset.py:166-177 (`get_position`, type 'single'): filled inclusive rectangle of 255."""
    h, w = shape
    x, y, bw, bh = box[:4]
    xa, xb = sorted((int(x), int(x + bw)))
    ya, yb = sorted((int(y), int(y + bh)))
    m = np.zeros((h, w), dtype=np.uint8)
    m[max(ya, 0): min(yb, h - 1) + 1, max(xa, 0): min(xb, w - 1) + 1] = 255
    return m


def rotate_bound_zero(image):
    """JointsDataset.py:179-201 with angle 0: identity, except that odd sizes pick up a half-pixel shift."""
    h, w = image.shape[:2]
    m = np.array([[1.0, 0.0, w / 2 - w // 2], [0.0, 1.0, h / 2 - h // 2]])
    return warp_affine_u8(image, m, (w, h))


def person_inputs(image, center, scale, box, image_size):
    """One person of one image -> (x [3,H,W] float32, pos_mask [1,H,W] float32)  (JointsDataset.py:296-331)."""
    trans = get_affine_transform(center, scale, image_size)
    crop = warp_affine_u8(image, trans, image_size)
    x = (crop.transpose(2, 0, 1).astype(np.float32) / np.float32(255)) - MEAN[:, None, None]
    x = x / STD[:, None, None]
    pm = resize_linear_u8(rotate_bound_zero(box_rectangle(image.shape[:2], box)), image_size)
    return x.astype(np.float32), (pm.astype(np.float32) / np.float32(255))[None]


def collate(images, annos, image_size):
    """List of images + per-image person lists -> (input [S,3,H,W], pos_mask [S,1,H,W], length)  (collater.py:14-26)."""
    xs, pms, length = [], [], []
    for img, people in zip(images, annos):
        for p in people:
            x, pm = person_inputs(img, p["center"], p["scale"], p["box"], image_size)
            xs.append(x)
            pms.append(pm)
        length.append(len(people))
    return np.stack(xs), np.stack(pms), length
