"""Parameter holders and small device programs shared by the model families (vanilla I2R-Net, TransPose-H first
stage, two-stage wrappers): the encoder parameter tree under the reference's key names and the
ConvTranspose2d(4, s=2, p=1)+BN+ReLU block as four phase problems."""
import torch
import torch.nn as nn

from .ops import ConvLayer
from .packing import deconv4x4s2_phase_taps, fold_bn


class EncoderLayerParams(nn.Module):
    """self_attn / linear1 / linear2 / norm1 / norm2 holder (reference interformer_pureMulti.py:169-180,
    transpose_h.py TransformerEncoderLayer, attention.py:37-58 -- identical parameter names)."""

    def __init__(self, d_model, nhead, dim_feedforward):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=0.1)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)


class EncoderParams(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([EncoderLayerParams(d_model, nhead, dim_feedforward) for _ in range(num_layers)])
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)


class DeconvProgram:
    """ConvTranspose2d(4, s=2, p=1) + BN + ReLU as four 2x2-tap phase problems in one grid."""

    def __init__(self, sd, conv_key, bn_key, device):
        w = sd[conv_key + ".weight"].float()          # [Cin, Cout, 4, 4]
        scale, bias = fold_bn(sd, bn_key, w.shape[1], conv_bias=sd.get(conv_key + ".bias"))
        self.phases = []
        for py in (0, 1):
            for px in (0, 1):
                mats, dys, dxs = deconv4x4s2_phase_taps(w, py, px)
                self.phases.append(((py, px), ConvLayer(mats, dys, dxs, scale, bias, relu=True, device=device)))
        self.cout = self.phases[0][1].cout      # padded channel count inside ops.channel_padding

    def run(self, r, x):
        nb, h, w, _ = x.shape
        cw = self.cout * (2 if self.phases[0][1].split else 1)      # split-operand tensors carry (hi | lo)
        out = torch.empty((nb, 2 * h, 2 * w, cw), dtype=torch.float16, device=x.device)
        r.conv_group([(L, x, dict(out=out, out_hw=(2 * h, 2 * w), out_mul=2, out_off=off)) for off, L in self.phases])
        return out




def make_deconv_stack(extra, with_bias):
    """nn.Sequential of NUM_DECONV_LAYERS x [ConvTranspose2d 4x4 s2 p1, BatchNorm2d, ReLU] (reference
    `_make_deconv_layer`, interformer.py:181-220 / interformer_2stage.py:296-319)."""
    nl, nf, nk = extra.NUM_DECONV_LAYERS, list(extra.NUM_DECONV_FILTERS), list(extra.NUM_DECONV_KERNELS)
    assert nl == len(nf), "ERROR: num_deconv_layers is different len(num_deconv_filters)"
    assert nl == len(nk), "ERROR: num_deconv_layers is different len(num_deconv_filters)"
    mods = []
    for i in range(nl):
        if nk[i] != 4:
            raise NotImplementedError("deconv kernel %d (shipped configs use 4)" % nk[i])
        mods += [nn.ConvTranspose2d(nf[i], nf[i], 4, 2, 1, 0, bias=with_bias), nn.BatchNorm2d(nf[i], momentum=0.1),
                 nn.ReLU(inplace=True)]
    return nn.Sequential(*mods)
