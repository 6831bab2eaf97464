"""`from utils.transforms import flip_back` -- lib/utils/transforms.py:16-30 on the device (csrc/postproc.cu).
CUDA tensor in -> CUDA tensor out (no host round trip; prefer `model.forward_flip`, which also fuses the averaging);
numpy in -> numpy out through the GPU, as the reference's `validate` calls it (lib/core/function.py:158)."""
import numpy as np
import torch

from i2r_b200 import postproc


def flip_back(output_flipped, matched_parts):
    """[N, K, H, W] heatmaps of mirrored inputs -> un-mirrored: width axis reversed, left/right joints swapped."""
    is_np = isinstance(output_flipped, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(output_flipped, dtype=np.float32)).cuda() if is_np else output_flipped
    assert t.dim() == 4, 'output_flipped should be [batch_size, num_joints, height, width]'
    perm = postproc.flip_permutation(matched_parts, t.shape[1], t.device).long()
    y = postproc.hflip(t.float().index_select(1, perm))
    return y.cpu().numpy() if is_np else y
