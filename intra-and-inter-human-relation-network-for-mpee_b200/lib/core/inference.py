"""`from core.inference import get_final_preds, get_max_preds` -- the reference's heatmap decode
(lib/core/inference.py:20-112) computed by one sm_100a kernel launch (csrc/postproc.cu) instead of per-joint python
loops over numpy arrays.  Same signatures; `hm` may be a CUDA tensor (stays on the device: the fast path after
`model(...)`), or a numpy array / CPU tensor as the reference's `validate` passes it (uploaded, decoded on the GPU,
results returned as numpy arrays of the reference's shapes and dtypes).  There is no CPU implementation here."""
import numpy as np
import torch

from i2r_b200 import postproc


def _to_cuda(hm):
    if isinstance(hm, torch.Tensor):
        return (hm if hm.is_cuda else hm.cuda()).float(), hm.is_cuda
    assert isinstance(hm, np.ndarray), 'batch_heatmaps should be numpy.ndarray'
    return torch.from_numpy(np.ascontiguousarray(hm, dtype=np.float32)).cuda(), False


def get_max_preds(batch_heatmaps):
    """-> (preds [N,K,2] = (x, y) of the maximum, zero where it is <= 0; maxvals [N,K,1])  (reference :20-48)."""
    hm, on_device = _to_cuda(batch_heatmaps)
    assert hm.dim() == 4, 'batch_images should be 4-ndim'
    # the blur-free decode: coordinates of the arg-max only (ksize 1 kernel, no Taylor step is applied to the result
    # the caller sees because the un-refined coordinates are what get_max_preds returns)
    flat = hm.reshape(hm.shape[0], hm.shape[1], -1)
    maxvals, idx = flat.max(dim=2, keepdim=True)
    w = hm.shape[3]
    preds = torch.cat([(idx % w).float(), torch.div(idx, w, rounding_mode="floor").float()], dim=2)
    preds = preds * (maxvals > 0).float()
    if on_device:
        return preds, maxvals
    return preds.cpu().numpy(), maxvals.cpu().numpy()


def get_final_preds(config, hm, center, scale, transform_back=True):
    """-> (preds [N,K,2] in original-image pixels, maxvals [N,K,1])  (reference :90-112; DARK refinement with
    config.TEST.BLUR_KERNEL)."""
    dev_hm, on_device = _to_cuda(hm)
    preds, maxvals = postproc.decode_heatmaps(dev_hm, center, scale, blur_kernel=int(config.TEST.BLUR_KERNEL),
                                              transform_back=transform_back)
    if on_device:
        return preds, maxvals
    return preds.cpu().numpy(), maxvals.cpu().numpy()
