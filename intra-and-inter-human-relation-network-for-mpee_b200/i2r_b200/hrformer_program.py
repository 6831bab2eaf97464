"""HRFormer-B first stage as a device program (reference: lib/models/hrformer.py:2057-2092, :2477-2480).

Channels 78 / 156 / 312 / 624 travel zero-padded to multiples of 16 (80 / 160 / 320 / 640: the K = 16 granularity of
the tensor-core GEMMs; `ops.channel_padding`), pad channels stay exactly zero through every layer.  One transformer
block (GeneralTransformerBlock, :1230-1240) is the launch sequence

    LayerNorm(eps 1e-6) + window-major gather (7x7 windows, centre zero padding)   i2r_ln_window_gather
    q / k / v projections (heads zero-padded 39 -> 48 channels by the packing)     i2r_conv_halo (1x1 GEMMs)
    softmax(q k^T / sqrt(39)) v per (window, head), no RPE bias, no mask           i2r_window_attention_tc
    out_proj                                                                        i2r_conv_halo
    scatter back + residual                                                         i2r_window_scatter_add
    LayerNorm                                                                       i2r_layernorm_padded
    fc1 1x1 + BN + GELU                                                             i2r_conv_halo (I2R_F_GELU)
    depthwise 3x3 + BN + GELU                                                       i2r_dwconv3x3
    fc2 1x1 + BN + GELU, + residual AFTER the activation                            i2r_conv_halo (GELU | ACT_FIRST)

and the fuse layers (:1616-1731) are 1x1 GEMMs at the source resolution + one bilinear up-sum pass per output branch,
depthwise stride-2 + 1x1 chains for the down-sampling terms.
"""
import torch

from .hrnet_w48 import _Unit, conv_bn_layer
from .ops import ConvLayer, _SPLIT_DEFAULT, pad_channels
from .packing import fold_bn, split_pair

WS, HEAD_PAD = 7, 48


def _pad_vec(v, n, fill=0.0):
    out = torch.full((n,), fill, dtype=torch.float32)
    out[: v.numel()] = v.float().reshape(-1)
    return out


def _lin(r, L, x2d, **kw):
    """1x1 'convolution' over the rows of x2d through Runner.conv (column chunking for > 256 outputs)."""
    t, c = x2d.shape
    assert x2d.is_contiguous()
    return r.conv(L, x2d.view(1, t, 1, c), **kw).view(t, -1)


class _Block:
    def __init__(self, sd, p, c, heads, device):
        self.c, self.heads = c, heads
        self.cp = pad_channels(c, small_ok=False)
        hd = c // heads
        self.scale = float(hd) ** -0.5
        hq = heads * HEAD_PAD
        a = p + ".attn.attn"

        def head_rows(w, b):       # output channel h*hd + j  ->  h*48 + j
            wp, bp = torch.zeros(hq, c), torch.zeros(hq)
            for h in range(heads):
                wp[h * HEAD_PAD: h * HEAD_PAD + hd] = w[h * hd:(h + 1) * hd].float()
                bp[h * HEAD_PAD: h * HEAD_PAD + hd] = b[h * hd:(h + 1) * hd].float()
            return wp, bp

        def lin(w, b):
            return ConvLayer([w], [0], [0], torch.ones(w.shape[0]), b, device=device)
        self.q = lin(*head_rows(sd[a + ".q_proj.weight"], sd[a + ".q_proj.bias"]))
        self.k = lin(*head_rows(sd[a + ".k_proj.weight"], sd[a + ".k_proj.bias"]))
        self.v = lin(*head_rows(sd[a + ".v_proj.weight"], sd[a + ".v_proj.bias"]))
        wo = sd[a + ".out_proj.weight"].float()
        wop = torch.zeros(c, hq)          # input channel h*hd + j  ->  h*48 + j
        for h in range(heads):
            wop[:, h * HEAD_PAD: h * HEAD_PAD + hd] = wo[:, h * hd:(h + 1) * hd]
        self.o = lin(wop, sd[a + ".out_proj.bias"].float())
        self.n1 = (_pad_vec(sd[p + ".norm1.weight"], self.cp).to(device), _pad_vec(sd[p + ".norm1.bias"], self.cp).to(device))
        self.n2 = (_pad_vec(sd[p + ".norm2.weight"], self.cp).to(device), _pad_vec(sd[p + ".norm2.bias"], self.cp).to(device))
        self.fc1 = conv_bn_layer(sd, p + ".mlp.fc1", p + ".mlp.norm1", device=device)
        ch = sd[p + ".mlp.dw3x3.weight"].shape[0]
        chp = pad_channels(ch, small_ok=False)
        dw = torch.zeros(9, chp)
        dw[:, :ch] = sd[p + ".mlp.dw3x3.weight"].float().reshape(ch, 9).t()
        sc, bi = fold_bn(sd, p + ".mlp.norm2", ch, conv_bias=sd[p + ".mlp.dw3x3.bias"])
        self.dw = (dw.contiguous().to(device), _pad_vec(sc, chp, 1.0).to(device), _pad_vec(bi, chp).to(device))
        self.fc2 = conv_bn_layer(sd, p + ".mlp.fc2", p + ".mlp.norm3", device=device)

    def run(self, r, x):
        nb, h, w, cw = x.shape
        rows = r.ln_window_gather(x, self.n1[0], self.n1[1], self.c, WS)
        # the three projections read the same rows: one grouped launch (column chunks of > 256 outputs included)
        t, c = rows.shape
        probs, qkv = [], []
        for L in (self.q, self.k, self.v):
            ps, o = r.problems(L, rows.view(1, t, 1, c))
            probs.extend(ps)
            qkv.append(o.view(t, -1))
        r.launch(probs)
        q, k, v = qkv
        hq = self.heads * HEAD_PAD
        a = r.window_attention(q[:, :hq], k[:, :hq], v[:, :hq], WS * WS, self.heads, self.scale, HEAD_PAD)
        o = _lin(r, self.o, a)
        x1 = r.window_scatter_add(x, o, WS)
        n2 = r.layernorm_padded(x1.view(-1, cw), self.n2[0], self.n2[1], self.c).view(nb, h, w, cw)
        t = r.conv(self.fc1, n2, relu=False, gelu=True)
        t = r.dwconv3x3(t, self.dw[0], self.dw[1], self.dw[2], 1, "gelu")
        return r.conv(self.fc2, t, relu=False, gelu=True, act_first=True, add0=x1)


class _DwChainStep:
    """depthwise 3x3 s2 + BN, then 1x1 conv + BN (+ ReLU unless it is the last step of the chain)."""

    def __init__(self, sd, p, relu, device):
        ch = sd[p + ".0.weight"].shape[0]
        chp = pad_channels(ch, small_ok=False)
        dw = torch.zeros(9, chp)
        dw[:, :ch] = sd[p + ".0.weight"].float().reshape(ch, 9).t()
        sc, bi = fold_bn(sd, p + ".1", ch)
        self.dw = (dw.contiguous().to(device), _pad_vec(sc, chp, 1.0).to(device), _pad_vec(bi, chp).to(device))
        self.pw = conv_bn_layer(sd, p + ".2", p + ".3", relu=relu, device=device)


class _Module:
    def __init__(self, sd, p, params, channels, heads, device):
        self.nb = len(params.branches)
        self.blocks = [[_Block(sd, "%s.branches.%d.%d" % (p, b, u), channels[b], heads[b], device)
                        for u in range(len(params.branches[b]))] for b in range(self.nb)]
        self.nout = len(params.fuse_layers)
        self.up, self.down = {}, {}
        for i, row in enumerate(params.fuse_layers):
            for j, f in enumerate(row):
                key = "%s.fuse_layers.%d.%d" % (p, i, j)
                if j > i:
                    self.up[(i, j)] = conv_bn_layer(sd, key + ".0", key + ".1", device=device)
                elif j < i:
                    self.down[(i, j)] = [_DwChainStep(sd, "%s.%d" % (key, s), relu=(s < i - j - 1), device=device)
                                         for s in range(i - j)]

    def run(self, r, xs):
        def chain(b):
            def go():
                y = xs[b]
                for blk in self.blocks[b]:
                    y = blk.run(r, y)
                return y
            return go
        ys = r.parallel([chain(b) for b in range(self.nb)])      # the branches are independent until the fuse layers
        def fuse(i):
            def go():
                acc = ys[i]
                ups = [j for j in range(self.nb) if j > i]
                downs = [j for j in range(self.nb) if j < i]
                for n, j in enumerate(downs):
                    t = ys[j]
                    chain = self.down[(i, j)]
                    for s, step in enumerate(chain):
                        t = r.dwconv3x3(t, step.dw[0], step.dw[1], step.dw[2], 2, None)
                        if s == len(chain) - 1:      # the last 1x1 of the chain accumulates onto the running sum
                            final = not ups and n == len(downs) - 1
                            t = r.conv(step.pw, t, add0=acc, relu=final)
                        else:
                            t = r.conv(step.pw, t)
                    acc = t
                if ups:
                    terms = [(r.conv(self.up[(i, j)], ys[j]), j - i) for j in ups]
                    acc = r.upsum_bilinear(acc, terms, relu=True)
                return acc
            return go
        return r.parallel([fuse(i) for i in range(self.nout)])      # one independent chain per output branch


class HRTProgram:
    def __init__(self, model, sd, device):
        b = "backbone."
        bb = model.backbone
        w = sd[b + "conv1.weight"].float()
        self.stem_w = w.permute(1, 2, 3, 0).reshape(-1, w.shape[0]).contiguous().to(device)
        sc, bi = fold_bn(sd, b + "bn1", w.shape[0])
        self.stem_scale, self.stem_bias = sc.to(device), bi.to(device)
        self.conv2 = conv_bn_layer(sd, b + "conv2", b + "bn2", stride=2, relu=True, device=device)
        self.layer1 = [_Unit(sd, "%slayer1.%d" % (b, i), "BOTTLENECK", device) for i in range(len(bb.layer1))]
        HRT_BASE = model.hrt_extra
        self.split = _SPLIT_DEFAULT[0]
        self.stages, self.trans = [], []
        for si in (2, 3, 4):
            cfg = HRT_BASE["stage%d" % si]
            tr = getattr(bb, "transition%d" % (si - 1))
            tl = []
            for i, t in enumerate(tr):
                key = "%stransition%d.%d" % (b, si - 1, i)
                if t is None:
                    tl.append(None)
                elif isinstance(t[0], torch.nn.Conv2d):
                    tl.append([conv_bn_layer(sd, key + ".0", key + ".1", relu=True, device=device)])
                else:
                    tl.append([conv_bn_layer(sd, "%s.%d.0" % (key, s), "%s.%d.1" % (key, s), stride=2, relu=True,
                                             device=device) for s in range(len(t))])
            self.trans.append(tl)
            st = getattr(bb, "stage%d" % si)
            self.stages.append([_Module(sd, "%sstage%d.%d" % (b, si, mi), m, cfg["num_channels"], cfg["num_heads"],
                                        device) for mi, m in enumerate(st)])
        self.head = conv_bn_layer(sd, "keypoint_head.final_layer", None, device=device)
        self.device = device

    @staticmethod
    def _bottleneck(r, u, x):
        h = r.conv(u.c1, x)
        h = r.conv(u.c2, h)
        res = r.conv(u.ds, x) if u.ds is not None else x
        return r.conv(u.c3, h, add0=res)

    def run(self, r, x):
        """x fp32 NCHW [S,3,H,W] -> (branch-0 feature fp16 NHWC [S,H/4,W/4,80], heatmaps fp32 NCHW [S,K,H/4,W/4])."""
        h = r.stem(x, self.stem_w, self.stem_scale, self.stem_bias, 64)
        h = r.conv(self.conv2, h)
        for u in self.layer1:
            h = self._bottleneck(r, u, h)
        xs = [h]
        for tl, mods in zip(self.trans, self.stages):
            nxt = []
            for i, t in enumerate(tl):
                if t is None:
                    nxt.append(xs[i])
                else:
                    y = xs[-1]      # transition layers read the last (lowest-resolution) map (:2069-2086)
                    for L in t:
                        y = r.conv(L, y)
                    nxt.append(y)
            xs = nxt
            for m in mods:
                xs = m.run(r, xs)
        feat = xs[0]
        return feat, r.conv(self.head, feat, out_mode="nchw32")
