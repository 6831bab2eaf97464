"""Post-norm Transformer encoder over ragged token sequences (inter-human stage; also the TransPose-H
intra stage).  Reference semantics: TransformerEncoderLayer.forward_post
(lib/models/interformer_pureMulti.py:182-213; lib/models/attention.py:61-82) on top of
nn.MultiheadAttention(nhead=1): q = k = (src+pos) W_qk + b, v = src W_v + b, q scaled by
head_dim**-0.5, softmax over the keys of the same image (padded persons never exist here: sequences
are stored ragged, which is exactly equivalent to key_padding_mask on zero-padded persons because
padded query rows are discarded by get_valid_output).
"""
import math

import torch

from .ops import ConvLayer, EncoderTailParams, pad_channels


def _linear_layer(weight, bias, relu=False, device="cuda"):
    return ConvLayer([weight.float()], [0], [0], torch.ones(weight.shape[0]), bias.float(), relu=relu, device=device)


class EncoderLayerProgram:
    def __init__(self, sd, prefix, d_model, device):
        w = sd[prefix + ".self_attn.in_proj_weight"].float()
        b = sd[prefix + ".self_attn.in_proj_bias"].float()
        d = d_model
        self.d = d
        self.qk = _linear_layer(w[: 2 * d], b[: 2 * d], device=device)
        self.v = _linear_layer(w[2 * d:], b[2 * d:], device=device)
        self.separate_qk = self.qk.split or pad_channels(d, small_ok=False) != d
        if self.separate_qk:
            # separate q / k projections: each output row is then one contiguous (hi | lo) pair, the operand
            # layout of the tcgen05 attention kernel
            self.q = _linear_layer(w[:d], b[:d], device=device)
            self.k = _linear_layer(w[d: 2 * d], b[d: 2 * d], device=device)
        self.out = _linear_layer(sd[prefix + ".self_attn.out_proj.weight"], sd[prefix + ".self_attn.out_proj.bias"],
                                 device=device)
        self.ff1 = _linear_layer(sd[prefix + ".linear1.weight"], sd[prefix + ".linear1.bias"], relu=True,
                                 device=device)
        self.ff2 = _linear_layer(sd[prefix + ".linear2.weight"], sd[prefix + ".linear2.bias"], device=device)
        self.tail = None
        if d == 96 and sd[prefix + ".linear1.weight"].shape[0] == 192:
            self.tail = EncoderTailParams(
                sd[prefix + ".self_attn.out_proj.weight"], sd[prefix + ".self_attn.out_proj.bias"],
                sd[prefix + ".linear1.weight"], sd[prefix + ".linear1.bias"], sd[prefix + ".linear2.weight"],
                sd[prefix + ".linear2.bias"], sd[prefix + ".norm1.weight"], sd[prefix + ".norm1.bias"],
                sd[prefix + ".norm2.weight"], sd[prefix + ".norm2.bias"], device=device, split=self.qk.split)
        dp = pad_channels(d, small_ok=False)

        def vec(key):      # LayerNorm parameters, zero-padded with the channels
            v = torch.zeros(dp)
            v[:d] = sd[prefix + key].float()
            return v.to(device)
        self.n1 = (vec(".norm1.weight"), vec(".norm1.bias"))
        self.n2 = (vec(".norm2.weight"), vec(".norm2.bias"))


class EncoderProgram:
    """`layers`-deep encoder; tokens [T, d] fp16, sequences delimited by cu_seqlens (int32, device)."""

    def __init__(self, sd, prefix, num_layers, d_model, nhead, device, normalize_before=False):
        if nhead != 1:
            raise NotImplementedError("N_HEAD=%d (all shipped configs use 1 head)" % nhead)
        if normalize_before:
            raise NotImplementedError("NORMALIZE_BEFORE=True (default False in every shipped config)")
        self.layers = [EncoderLayerProgram(sd, "%s.layers.%d" % (prefix, i), d_model, device)
                       for i in range(num_layers)]
        self.d = pad_channels(d_model, small_ok=False)      # channels per token as stored (zero-padded to 16)
        self.d_real = d_model
        self.scale = 1.0 / math.sqrt(d_model // nhead)

    def run(self, r, src, pos, cu_seqlens, max_seqlen, attn_tap=None):
        """src / pos: [T, d] fp16, or split tensors [T, 2d] (hi | lo) when the layers were built in split-operand mode.
        attn_tap(layer, q [T,d] fp32, k [T,d] fp32, scale): visualisation aid -- called with the projected queries / keys
        of every layer when forward hooks are registered on the `self_attn` holders (module_base.attention_hooks)."""
        d = self.d
        split = self.layers[0].qk.split
        sp = r.add(src, pos) if pos is not None else src
        for li, L in enumerate(self.layers):
            # V is written transposed (channel-major): the K-major B operand of the P.V product
            pv, vt = r.linear_problem(L.v, src, out_mode="t16")
            if L.separate_qk:     # q, k: [T, (hi | lo)] (or [T, d_pad]); vt: [(hi rows | lo rows), T]
                pq, q = r.linear_problem(L.q, sp)
                pk, k = r.linear_problem(L.k, sp)
                r.launch([pq, pk, pv])
                wq = 2 * d if split else d
                if attn_tap is not None:
                    q2, k2 = q.view(-1, wq).float(), k.view(-1, wq).float()
                    if split:
                        q2, k2 = q2[:, :d] + q2[:, d:], k2[:, :d] + k2[:, d:]
                    attn_tap(li, q2[:, :self.d_real], k2[:, :self.d_real], self.scale)
                a = r.attention_tc(q.view(-1, wq), k.view(-1, wq), vt, cu_seqlens, max_seqlen, self.scale,
                                   split=split)
            else:
                pq, qk = r.linear_problem(L.qk, sp)
                r.launch([pq, pv])
                qk = qk.view(-1, 2 * d)
                if attn_tap is not None:
                    attn_tap(li, qk[:, :d].float(), qk[:, d:].float(), self.scale)
                a = r.attention_tc(qk[:, :d], qk[:, d:], vt, cu_seqlens, max_seqlen, self.scale)
            last = li == len(self.layers) - 1
            if L.tail is not None:
                src, sp2 = r.encoder_tail(L.tail, a, src, pos=None if (last or pos is None) else pos)
                sp = sp2 if sp2 is not None else src
                continue
            x1 = r.linear(L.out, a, add0=src)
            if self.d != self.d_real:      # zero-padded channels (d_model 78 -> 80): statistics over the real ones
                if pos is not None:
                    raise NotImplementedError("position embedding with a padded d_model")
                s1 = r.layernorm_padded(x1, L.n1[0], L.n1[1], self.d_real, eps=1e-5)
                h = r.linear(L.ff1, s1)
                x2 = r.linear(L.ff2, h, add0=s1)
                src = sp = r.layernorm_padded(x2, L.n2[0], L.n2[1], self.d_real, eps=1e-5)
                continue
            s1, _ = r.layernorm(x1, L.n1[0], L.n1[1])
            h = r.linear(L.ff1, s1)
            x2 = r.linear(L.ff2, h, add0=s1)
            last = li == len(self.layers) - 1
            src, sp2 = r.layernorm(x2, L.n2[0], L.n2[1], pos=None if (last or pos is None) else pos)
            sp = sp2 if sp2 is not None else src
        return src
