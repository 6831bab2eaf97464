"""Position embeddings of the path.

* `sine_table`  -- fixed 2-D sine embedding [h*w, 1, d] (reference: interformer_pureMulti.py:516-541,
  transpose_h.py `_make_sine_position_embedding`): y/x cumulative positions normalised to 2*pi,
  temperature 10000, sin on even / cos on odd feature indices, y-half then x-half of the channels.
* `MaskEmbedParams` / `MaskEmbedProgram` -- the multi-person embedding computed from each person's
  box mask (reference: lib/models/position_embedding.py:6-117, mode 'conv'):
  conv3x3 s2 1->64 + BN + ReLU, conv3x3 s2 64->d + BN + ReLU, then log2(W/4 / trans_w) max-pools.
"""
import math

import torch
import torch.nn as nn

from .hrnet_w48 import conv_bn_layer
from .packing import fold_bn


def sine_table(h, w, d_model, temperature=10000.0, scale=2 * math.pi):
    half = d_model // 2
    eps = 1e-6
    ys = torch.arange(1, h + 1, dtype=torch.float32).view(h, 1).expand(h, w)
    xs = torch.arange(1, w + 1, dtype=torch.float32).view(1, w).expand(h, w)
    ys = ys / (float(h) + eps) * scale
    xs = xs / (float(w) + eps) * scale
    idx = torch.arange(half, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(idx, 2, rounding_mode="floor") / half)

    def interleave(p):                       # [h, w, half]: sin on even, cos on odd feature indices
        return torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=3).flatten(2)

    py = interleave(ys[..., None] / dim_t)
    px = interleave(xs[..., None] / dim_t)
    pos = torch.cat((py, px), dim=2)         # [h, w, d]
    return pos.reshape(h * w, 1, d_model)


class MaskEmbedParams(nn.Module):
    """Parameter holder under the reference's names (`conv1/bn1/conv2/bn2`, or `fc`, or `conv_pre/res/conv_end`)."""

    def __init__(self, trans_size, d_model=96, mode="conv", vec_dim=None):
        super().__init__()
        self.trans_size, self.d_model, self.mode = list(trans_size), d_model, mode
        if mode == "conv":
            self.conv1 = nn.Conv2d(1, 64, 3, 2, 1, bias=False)
            self.bn1 = nn.BatchNorm2d(64, momentum=0.1)
            self.conv2 = nn.Conv2d(64, d_model, 3, 2, 1, bias=False)
            self.bn2 = nn.BatchNorm2d(d_model, momentum=0.1)
        elif mode == "cat_vec":
            self.fc = nn.Linear(self.trans_size[0] * self.trans_size[1], vec_dim)
        elif mode == "res":
            from torchvision.models import resnet18
            self.conv_pre = nn.Conv2d(1, 3, 3, 1, 1, bias=False)
            self.res = nn.Sequential(*list(resnet18(weights=None).children())[:5])
            self.conv_end = nn.Conv2d(64, d_model, 3, 1, 1, bias=False)
        elif mode != "sine":
            raise ValueError("unknown MULTI_POS_EMBEDDING %r" % mode)


class MaskEmbedResProgram:
    """Mode 'res' (reference position_embedding.py:14-18, :90-108): conv_pre + resnet18 stem as one kernel
    (csrc/mask_res.cu), max-pool, resnet18.layer1 (two BasicBlocks) and conv_end on the generic conv kernels, then the
    max-pools down to TRANS_SIZE."""

    def __init__(self, sd, prefix, device):
        from .hrnet_w48 import _Unit
        self.w_pre = sd[prefix + ".conv_pre.weight"].float().reshape(3, 9).contiguous().to(device)
        w1 = sd[prefix + ".res.0.weight"].float()                      # [64,3,7,7] -> [147,64], k = (c*7+ky)*7+kx
        self.w1 = w1.permute(1, 2, 3, 0).reshape(147, 64).contiguous().to(device)
        sc, bi = fold_bn(sd, prefix + ".res.1", 64)
        self.scale, self.bias = sc.to(device), bi.to(device)
        self.blocks = [_Unit(sd, "%s.res.4.%d" % (prefix, b), "BASIC", device) for b in (0, 1)]
        self.conv_end = conv_bn_layer(sd, prefix + ".conv_end", None, device=device)

    def run(self, r, pos_mask, trans_hw):
        y = r.mask_res_stem(pos_mask, self.w_pre, self.w1, self.scale, self.bias)
        y = r.maxpool(y)
        for u in self.blocks:
            h = r.conv(u.c1, y)
            y = r.conv(u.c2, h, add0=y)
        y = r.conv(self.conv_end, y)
        for _ in range(int(math.log(y.shape[2] // trans_hw[1], 2))):
            y = r.maxpool(y)
        return y


def build_mask_embed_program(mode, sd, prefix, device):
    if mode == "conv":
        return MaskEmbedProgram(sd, prefix, device)
    if mode == "res":
        return MaskEmbedResProgram(sd, prefix, device)
    raise NotImplementedError("MULTI_POS_EMBEDDING=%r with USE_MULTI_POS (kernels exist for 'conv' and 'res')" % mode)


class MaskEmbedProgram:
    def __init__(self, sd, prefix, device):
        w = sd[prefix + ".conv1.weight"].float()           # [64,1,3,3] -> [9,64]
        self.c1_w = w.permute(1, 2, 3, 0).reshape(-1, w.shape[0]).contiguous().to(device)
        sc, bi = fold_bn(sd, prefix + ".bn1", w.shape[0])
        self.c1_scale, self.c1_bias, self.c1_out = sc.to(device), bi.to(device), w.shape[0]
        self.c2 = conv_bn_layer(sd, prefix + ".conv2", prefix + ".bn2", stride=2, relu=True, device=device)

    def run(self, r, pos_mask, trans_hw):
        """pos_mask: fp32 [S,1,H,W] -> fp16 [S, th, tw, d]."""
        y = r.stem(pos_mask, self.c1_w, self.c1_scale, self.c1_bias, self.c1_out)
        y = r.conv(self.c2, y)
        for _ in range(int(math.log(y.shape[2] // trans_hw[1], 2))):
            y = r.maxpool(y)
        return y
