"""TransPose-H first stage as a device program: HRNet-W48-S -> `reduce` 1x1 on branch HRNET_RES_LAYER ->
post-norm intra-human encoder over the (H/4 * W/4) tokens of EACH crop with a fixed position table ->
(feature map, first-stage heatmaps).  Reference: lib/models/transpose_h.py:623-655.

Sequences are the crops themselves (cu_seqlens = multiples of H/4*W/4), so the attention kernel that serves the
ragged inter-human stage serves this one unchanged; the [L,1,d] position parameter is tiled over the crops once
per batch size.
"""
import torch

from .encoder import EncoderProgram
from .hrnet_w48 import BackboneProgram, conv_bn_layer
from .ops import _SPLIT_DEFAULT
from .packing import split_pair


class FirstStageProgram:
    def __init__(self, model, sd, device):
        c = model._cfg
        self.res_layer = c["res_layer"]
        self.backbone = BackboneProgram(model, sd, device)
        self.reduce = conv_bn_layer(sd, "reduce", None, device=device)
        self.encoder = EncoderProgram(sd, "global_encoder", c["layers"], c["d_model"], c["nhead"], device)
        if c["final_k"] != 1:
            raise NotImplementedError("FINAL_CONV_KERNEL=3")
        self.head = conv_bn_layer(sd, "final_layer", None, device=device)
        pe = sd.get("pos_embedding")
        self.split = _SPLIT_DEFAULT[0]
        if pe is None:
            self.pos_table = None
        else:
            pe = pe.float().reshape(pe.shape[0], -1)
            self.pos_table = (split_pair(pe) if self.split else pe.half()).to(device)
        self.device = device
        self._pos, self._cu = {}, {}

    def _tiled(self, crops, tokens):
        if crops not in self._cu:
            self._cu[crops] = (torch.arange(crops + 1, dtype=torch.int32) * tokens).to(self.device)
            self._pos[crops] = None if self.pos_table is None else self.pos_table.repeat(crops, 1).contiguous()
        return self._pos[crops], self._cu[crops]

    def run(self, r, x):
        """x fp32 NCHW [S,3,H,W] -> (feat fp16 NHWC [S,h,w,d], heatmaps fp32 NCHW [S,K,h,w])."""
        feats = self.backbone.run(r, x)
        f = r.conv(self.reduce, feats[self.res_layer])
        s, h, w, d = f.shape      # d = 2 * d_model in split-operand mode (hi | lo)
        if self.pos_table is not None and self.pos_table.shape[0] != h * w:
            raise ValueError("pos_embedding holds %d tokens, the feature map %d" % (self.pos_table.shape[0], h * w))
        pos, cu = self._tiled(s, h * w)
        y = self.encoder.run(r, f.view(s * h * w, d), pos, cu, h * w)
        feat = y.view(s, h, w, d)
        return feat, r.conv(self.head, feat, out_mode="nchw32")
