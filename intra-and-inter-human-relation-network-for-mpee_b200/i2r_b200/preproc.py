"""Test-time input pipeline on the device: what `JointsDataset.__getitem__` (is_train False,
lib/dataset/JointsDataset.py:205-340) and `collater(0)` (lib/dataset/collater.py:14-26, :175-181) hand to the forward --
`(input [S,3,H,W], pos_mask [S,1,H,W], length)` -- computed from the decoded uint8 images and the person boxes by two
kernels per image (csrc/preproc.cu) instead of cv2 calls and torch ops in DataLoader workers.

The 2x3 matrices are host arithmetic (a 6x6 solve per person, as cv2.getAffineTransform); pixels never touch the host
after the image upload."""
import ctypes

import numpy as np
import torch

from . import capi

IMAGENET_MEAN = (0.485, 0.456, 0.406)      # tools/test.py:126-128
IMAGENET_STD = (0.229, 0.224, 0.225)


def get_affine_transform(center, scale, output_size):
    """lib/utils/transforms.py:58-92 with rot = 0, shift = 0, inv = 0: float32 point triplets, double 6x6 solve."""
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
    """The dst -> src matrix cv2.warpAffine computes from a forward matrix (same operations, same order)."""
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


def box_corners(box):
    """Inclusive corners cv2.rectangle fills for `get_position` (JointsDataset.py:169-170)."""
    x, y, w, h = (float(v) for v in box[:4])
    xa, xb = sorted((int(x), int(x + w)))
    ya, yb = sorted((int(y), int(y + h)))
    return xa, ya, xb, yb


class GpuCropper:
    """`cropper(images, annos)` -> (input, pos_mask, length) on `device`.

    images: list of uint8 RGB arrays / tensors [H, W, 3] (what cv2.imread + COLOR_BGR2RGB yields, :215-222);
    annos: per image, a list of dicts with 'center' [2], 'scale' [2], 'box' [x, y, w, h] (the db records, :266-275)."""

    def __init__(self, image_size, device="cuda", mean=IMAGENET_MEAN, std=IMAGENET_STD):
        self.image_size = (int(image_size[0]), int(image_size[1]))      # (W, H) as cfg.MODEL.IMAGE_SIZE
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise capi.I2RError("GpuCropper runs on a CUDA device only (there is no CPU path)")
        self.lib = capi.load()
        self._mean = (ctypes.c_float * 3)(*mean)
        self._std = (ctypes.c_float * 3)(*std)

    def __call__(self, images, annos):
        ow, oh = self.image_size
        length = [len(a) for a in annos]
        s = sum(length)
        x = torch.empty((s, 3, oh, ow), dtype=torch.float32, device=self.device)
        pm = torch.empty((s, 1, oh, ow), dtype=torch.float32, device=self.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        keep, at = [], 0
        for img, people in zip(images, annos):
            n = len(people)
            if n == 0:
                continue
            t = torch.as_tensor(img)
            assert t.dtype == torch.uint8 and t.dim() == 3 and t.shape[2] == 3, "images are uint8 [H, W, 3] RGB"
            ih, iw = int(t.shape[0]), int(t.shape[1])
            dimg = t.contiguous().to(self.device, non_blocking=True)
            inv = np.stack([invert_affine(get_affine_transform(p["center"], p["scale"], self.image_size)) for p in people])
            rect = np.asarray([box_corners(p["box"]) for p in people], dtype=np.int32)
            dinv = torch.from_numpy(inv).to(self.device, non_blocking=True)
            drect = torch.from_numpy(rect).to(self.device, non_blocking=True)
            capi.check(self.lib.i2r_crop_persons(dimg.data_ptr(), ih, iw, dinv.data_ptr(), n, oh, ow, self._mean, self._std,
                                                 x[at:at + n].data_ptr(), stream), "i2r_crop_persons")
            capi.check(self.lib.i2r_box_masks(drect.data_ptr(), n, ih, iw, oh, ow, pm[at:at + n].data_ptr(), stream),
                       "i2r_box_masks")
            keep += [dimg, dinv, drect]
            at += n
        self._keep = keep      # alive until the next call (the launches above are asynchronous)
        return x, pm, length
