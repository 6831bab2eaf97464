"""Device-side post-processing of the heatmaps: the flip test without its numpy round trips
(lib/core/function.py:142-162, lib/utils/transforms.py:16-30) and `get_final_preds` (lib/core/inference.py:20-112) as
one kernel launch instead of per-joint python loops.  Thin wrappers over libi2r_sm100.so; CUDA tensors only."""
import ctypes

import torch

from . import capi


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise capi.I2RError("%s runs on CUDA tensors only (there is no CPU path)" % what)


def hflip(x):
    """np.flip(x, -1) for a contiguous fp32 CUDA tensor (function.py:145-146)."""
    _need_cuda(x, "hflip")
    x = x.contiguous()
    assert x.dtype == torch.float32
    y = torch.empty_like(x)
    capi.check(capi.load().i2r_hflip_f32(x.data_ptr(), y.data_ptr(), x.numel() // x.shape[-1], x.shape[-1], _stream()),
               "i2r_hflip_f32")
    return y


_PERMS = {}


def flip_permutation(flip_pairs, num_joints, device):
    key = (tuple(tuple(int(v) for v in p) for p in flip_pairs), int(num_joints), str(device))
    if key not in _PERMS:
        perm = list(range(num_joints))
        for a, b in key[0]:
            perm[a], perm[b] = perm[b], perm[a]
        _PERMS[key] = torch.tensor(perm, dtype=torch.int32).to(device)
    return _PERMS[key]


def flip_merge(out, out_flipped, flip_pairs):
    """(out + flip_back(out_flipped, flip_pairs)) * 0.5 on fp32 [S,K,h,w] CUDA heatmaps (function.py:158-162)."""
    _need_cuda(out, "flip_merge")
    _need_cuda(out_flipped, "flip_merge")
    out, out_flipped = out.contiguous(), out_flipped.contiguous()
    assert out.shape == out_flipped.shape and out.dim() == 4 and out.dtype == torch.float32
    s, k, h, w = out.shape
    y = torch.empty_like(out)
    perm = flip_permutation(flip_pairs, k, out.device)
    capi.check(capi.load().i2r_flip_merge(out.data_ptr(), out_flipped.data_ptr(), y.data_ptr(), s, k, h, w,
                                          perm.data_ptr(), _stream()), "i2r_flip_merge")
    return y


def decode_heatmaps(hm, center=None, scale=None, blur_kernel=11, transform_back=True):
    """get_final_preds on the device.  hm: fp32 CUDA [S,K,H,W]; center / scale: [S,2] (any device / numpy).  Returns
    (preds [S,K,2], maxvals [S,K,1]) fp32 CUDA tensors."""
    _need_cuda(hm, "decode_heatmaps")
    hm = hm.contiguous()
    assert hm.dim() == 4 and hm.dtype == torch.float32
    s, k, h, w = hm.shape
    c = sc = None
    if transform_back:
        c = torch.as_tensor(center, dtype=torch.float32).to(hm.device).contiguous()
        sc = torch.as_tensor(scale, dtype=torch.float32).to(hm.device).contiguous()
        assert c.shape == (s, 2) and sc.shape == (s, 2)
    preds = torch.empty((s, k, 2), dtype=torch.float32, device=hm.device)
    maxvals = torch.empty((s, k, 1), dtype=torch.float32, device=hm.device)
    capi.check(capi.load().i2r_decode_heatmaps(hm.data_ptr(), s, k, h, w, c.data_ptr() if c is not None else None,
                                               sc.data_ptr() if sc is not None else None, int(blur_kernel),
                                               int(bool(transform_back)), preds.data_ptr(), maxvals.data_ptr(),
                                               _stream()), "i2r_decode_heatmaps")
    return preds, maxvals
