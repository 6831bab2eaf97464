"""TEST INFRASTRUCTURE: a torch-CPU interpreter of the C-ABI op semantics (i2r_conv_problem etc.).

It executes the *same* problem structs, packed weights and launch sequence the product path sends to
libi2r_sm100.so, with fp16 storage and fp32 accumulation, so the host logic (BN folding, weight
packing, tap tables, fuse wiring, deconv phases, ragged attention, graph of launches) can be checked
against the golden vectors without a GPU.  It is never used by the product path.
"""
import torch
import torch.nn.functional as F

from i2r_b200 import capi
from i2r_b200.ops import Runner
from i2r_b200.packing import KC, ceil_to, merge_pair, split_pair, unpack_taps


class EmuRunner(Runner):
    def __init__(self):
        self.impl = -1
        self.use_tma = True     # exercise the N-split logic of Runner.problems
        self.timing = None
        self.launches = 0
        self.split = False
        self.device = torch.device("cpu")
        self._chain, self.chain_enabled, self._in_parallel, self.chains = None, False, 0, 0   # no chained launches here

    def launch(self, problems):
        for p in problems:
            self._run_problem(p)
        self.launches += 1

    @staticmethod
    def _run_problem(p):
        x, L, add0, add1, out = p._keep
        nb, hs, ws, _ = x.shape
        sh = p.in_shift
        cin, npad, cout = p.Cin, p.Npad, p.Cout
        split = bool(p.flags & capi.F_SPLIT)
        if split:
            # x = (hi | lo); product = (hi + lo) * W_hi + hi * W_lo with W = [W_hi | W_hi | W_lo] along K
            kp = ceil_to(cin, KC)
            wv = unpack_taps(L.w, 3 * kp)
            w = wv[:, :, :cin]
            w_lo = wv[:, :, 2 * kp:2 * kp + cin]
            assert torch.equal(wv[:, :, kp:kp + cin], w)
        else:
            w = unpack_taps(L.w, cin)                         # [ntaps, npad, cin]
        oy = torch.arange(p.OH).view(-1, 1).expand(p.OH, p.OW)
        ox = torch.arange(p.OW).view(1, -1).expand(p.OH, p.OW)
        acc = torch.zeros(nb, p.OH, p.OW, npad)
        xf = x[..., :cin].float()
        if split:
            x_hi = xf
            xf = xf + x[..., cin:2 * cin].float()
        for t in range(p.ntaps):
            wt = w[t]
            iy = oy * p.stride + int(p.dy[t])
            ix = ox * p.stride + int(p.dx[t])
            ok = (iy >= 0) & (iy < p.IH) & (ix >= 0) & (ix < p.IW)
            sy = (iy.clamp(0, p.IH - 1) >> sh)
            sx = (ix.clamp(0, p.IW - 1) >> sh)
            a = xf[:, sy, sx, :] * ok[None, :, :, None]
            acc += a @ wt.t()
            if split:
                acc += (x_hi[:, sy, sx, :] * ok[None, :, :, None]) @ w_lo[t].t()
        v = acc[..., :cout] * L.scale[:cout] + L.bias[:cout]
        fy = oy * p.out_mul + p.out_offy
        fx = ox * p.out_mul + p.out_offx
        lo_off = int(p.pair_lo_offset)

        def lo_view(t):     # the lo-half slice that sits lo_off channels after a hi-half slice of a pair tensor
            return t.as_strided(t.shape, t.stride(), t.storage_offset() + lo_off)

        def act(t):
            if p.flags & capi.F_GELU:
                return F.gelu(t)
            return F.relu(t) if p.flags & capi.F_RELU else t
        if p.flags & capi.F_ACT_FIRST:
            v = act(v)
        for a, s in ((add0, p.add0_shift), (add1, p.add1_shift)):
            if a is not None:
                assert tuple(a.shape) == (nb, p.OHf >> s, p.OWf >> s, 2 * cout if (split and not lo_off) else cout)
                assert a.stride(2) == p.add_pix_stride
                if lo_off:
                    af = a.float() + lo_view(a).float()
                else:
                    af = merge_pair(a) if split else a.float()
                v = v + af[:, fy >> s, fx >> s, :]
        if not (p.flags & capi.F_ACT_FIRST):
            v = act(v)
        if p.flags & capi.F_OUT_NCHW_F32:
            out[:, :, fy, fx] = v.permute(0, 3, 1, 2)
        elif p.flags & capi.F_OUT_F32:
            out[:, fy, fx, :] = v
        elif p.flags & capi.F_OUT_T16:
            vv = split_pair(v) if split else v.half()
            out.copy_(vv.reshape(-1, vv.shape[-1]).t())
        elif split and lo_off:
            pair = split_pair(v)
            out[:, fy, fx, :] = pair[..., :cout]
            lo_view(out)[:, fy, fx, :] = pair[..., cout:]
        elif split:
            out[:, fy, fx, :] = split_pair(v)
        else:
            out[:, fy, fx, :] = v.half()

    def _rd(self, x):
        return merge_pair(x) if self.split else x.float()

    def _wr(self, v):
        return split_pair(v) if self.split else v.half()

    def stem(self, x, w, scale, bias, cout):
        cin = x.shape[1]
        wt = w.reshape(cin, 3, 3, cout).permute(3, 0, 1, 2)
        y = F.conv2d(x, wt, None, 2, 1) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
        self.launches += 1
        return self._wr(F.relu(y).permute(0, 2, 3, 1).contiguous())

    def mask_res_stem(self, mask, w_pre, w1, scale, bias):
        pre = F.conv2d(mask, w_pre.reshape(3, 1, 3, 3), None, 1, 1)
        y = F.conv2d(pre, w1.reshape(3, 7, 7, 64).permute(3, 0, 1, 2), None, 2, 3)
        y = y * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
        self.launches += 1
        return self._wr(F.relu(y).permute(0, 2, 3, 1).contiguous())

    def maxpool(self, x):
        self.launches += 1
        return self._wr(F.max_pool2d(self._rd(x).permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).contiguous())

    def layernorm(self, x2d, gamma, beta, eps=1e-5, pos=None):
        xf = self._rd(x2d)
        y = F.layer_norm(xf, (xf.shape[1],), gamma, beta, eps)
        self.launches += 1
        return self._wr(y), (self._wr(y + self._rd(pos)) if pos is not None else None)

    def add(self, a, b):
        self.launches += 1
        return self._wr(self._rd(a) + self._rd(b))

    def upsum(self, x0, t1, shift1, t2=None, shift2=0, relu=True):
        nb, h, w, _ = x0.shape
        ys = torch.arange(h).view(-1, 1).expand(h, w)
        xs = torch.arange(w).view(1, -1).expand(h, w)
        v = self._rd(x0) + self._rd(t1)[:, ys >> shift1, xs >> shift1, :]
        if t2 is not None:
            v = v + self._rd(t2)[:, ys >> shift2, xs >> shift2, :]
        self.launches += 1
        return self._wr(F.relu(v) if relu else v)

    def attention(self, q, k, v, cu_seqlens, max_seqlen, scale, lo=None):
        t, d = q.shape
        if lo is not None:      # split operands: lo halves sit `lo` elements after the hi views in the same rows
            def lo_view(h, off):
                return h.as_strided(h.shape, h.stride(), h.storage_offset() + off)
            qh, kh, vh = q.float(), k.float(), v.float()
            ql, kl, vl = lo_view(q, lo[0]).float(), lo_view(k, lo[1]).float(), lo_view(v, lo[2]).float()
        out = torch.empty(t, 2 * d if lo is not None else d, dtype=torch.float16)
        cu = cu_seqlens.tolist()
        for a, b in zip(cu[:-1], cu[1:]):
            assert b - a <= max_seqlen
            if lo is None:
                s = torch.softmax(q[a:b].float() @ k[a:b].float().t() * scale, dim=-1)
                out[a:b] = (s.half().float() @ v[a:b].float()).half()
            else:
                sc = (qh[a:b] + ql[a:b]) @ kh[a:b].t() + qh[a:b] @ kl[a:b].t()      # three-term score product
                s = torch.softmax(sc * scale, dim=-1).half().float()                # probabilities stay fp16
                out[a:b] = split_pair(s @ (vh[a:b] + vl[a:b]))
        self.launches += 1
        return out

    def attention_tc(self, q, k, vt, cu_seqlens, max_seqlen, scale, split=False):
        t, w = q.shape
        d = w // 2 if split else w
        v = vt.t()
        out = torch.empty(t, w, dtype=torch.float16)
        cu = cu_seqlens.tolist()
        for a, b in zip(cu[:-1], cu[1:]):
            assert b - a <= max_seqlen
            if not split:
                s = torch.softmax(q[a:b].float() @ k[a:b].float().t() * scale, dim=-1)
                out[a:b] = (s.half().float() @ v[a:b].float()).half()
            else:
                qh, ql = q[a:b, :d].float(), q[a:b, d:].float()
                kh, kl = k[a:b, :d].float(), k[a:b, d:].float()
                sc = (qh + ql) @ kh.t() + qh @ kl.t()                               # three-term score product
                s = torch.softmax(sc * scale, dim=-1).half().float()                # probabilities stay fp16
                out[a:b] = split_pair(s @ (v[a:b, :d].float() + v[a:b, d:].float()))
        self.launches += 1
        return out

    def encoder_tail(self, tail, attn, src, pos=None, eps=1e-5):
        w_out, b_out, w1, b1, w2, b2, g1, be1, g2, be2 = tail.host

        def q(m):       # operand precision of the kernel: fp16, or the (hi + lo) pair in split mode
            hi = m.half().float()
            return hi + (m - hi).half().float() if tail.split else hi

        def act(v):     # activations handed from an epilogue to the next GEMM through shared memory
            return merge_pair(split_pair(v)) if tail.split else v.half().float()
        a, s0 = self._rd(attn), self._rd(src)
        s1 = act(F.layer_norm(a @ q(w_out).t() + b_out + s0, (96,), g1, be1, eps))
        h = act(F.relu(s1 @ q(w1).t() + b1))
        y = F.layer_norm(h @ q(w2).t() + b2 + s1, (96,), g2, be2, eps)
        self.launches += 1
        return self._wr(y), (self._wr(y + self._rd(pos)) if pos is not None else None)

    # ------------------------------------------------------------------ HRFormer-B building blocks
    def dwconv3x3(self, x, w, scale, bias, stride=1, act=None):
        c = w.shape[1]
        xf = self._rd(x).permute(0, 3, 1, 2)
        wt = w.t().reshape(c, 1, 3, 3)
        y = F.conv2d(xf, wt, None, stride, 1, 1, c) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
        y = {"gelu": F.gelu, "relu": F.relu, None: lambda t: t, "none": lambda t: t}[act](y)
        self.launches += 1
        return self._wr(y.permute(0, 2, 3, 1).contiguous())

    def upsum_bilinear(self, x0, terms, relu=True):
        v = self._rd(x0).permute(0, 3, 1, 2)
        for t, s in terms:
            v = v + F.interpolate(self._rd(t).permute(0, 3, 1, 2), scale_factor=2 ** s, mode="bilinear",
                                  align_corners=False)
        v = F.relu(v) if relu else v
        self.launches += 1
        return self._wr(v.permute(0, 2, 3, 1).contiguous())

    def _ln_padded(self, xf, gamma, beta, c_real, eps):
        y = torch.zeros_like(xf)
        y[..., :c_real] = F.layer_norm(xf[..., :c_real], (c_real,), gamma[:c_real], beta[:c_real], eps)
        return y

    def layernorm_padded(self, x2d, gamma, beta, c_real, eps=1e-6):
        self.launches += 1
        return self._wr(self._ln_padded(self._rd(x2d), gamma, beta, c_real, eps))

    @staticmethod
    def _win_geom(h, w, ws):
        ph, pw = (-h) % ws, (-w) % ws
        return ph, pw, h + ph, w + pw

    def ln_window_gather(self, x, gamma, beta, c_real, ws=7, eps=1e-6):
        nb, h, w, _ = x.shape
        ph, pw, hp, wp = self._win_geom(h, w, ws)
        ln = self._ln_padded(self._rd(x), gamma, beta, c_real, eps)
        xp = F.pad(ln, (0, 0, pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
        c = xp.shape[-1]
        rows = xp.view(nb, hp // ws, ws, wp // ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(-1, c)
        self.launches += 1
        return self._wr(rows.contiguous())

    def window_scatter_add(self, x, a, ws=7):
        nb, h, w, _ = x.shape
        ph, pw, hp, wp = self._win_geom(h, w, ws)
        af = self._rd(a)
        c = af.shape[-1]
        m = af.view(nb, hp // ws, wp // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(nb, hp, wp, c)
        m = m[:, ph // 2: ph // 2 + h, pw // 2: pw // 2 + w, :]
        self.launches += 1
        return self._wr(self._rd(x) + m)

    def window_attention(self, q, k, v, win_len, heads, scale, head_pad=48):
        t, cq = q.shape

        def val(m):
            if not self.split:
                return m.float()
            lo = m.as_strided(m.shape, m.stride(), m.storage_offset() + cq)
            return m.float(), lo.float()
        nwin = t // win_len
        if self.split:
            (qh, ql), (kh, kl), (vh, vl) = val(q), val(k), val(v)
            shp = lambda m: m.reshape(nwin, win_len, heads, head_pad).permute(0, 2, 1, 3)
            sc = shp(qh + ql) @ shp(kh).transpose(-1, -2) + shp(qh) @ shp(kl).transpose(-1, -2)
            p = torch.softmax(sc * scale, dim=-1).half().float()
            o = (p @ shp(vh + vl)).permute(0, 2, 1, 3).reshape(t, cq)
            out = split_pair(o)
        else:
            shp = lambda m: m.float().reshape(nwin, win_len, heads, head_pad).permute(0, 2, 1, 3)
            p = torch.softmax(shp(q) @ shp(k).transpose(-1, -2) * scale, dim=-1).half().float()
            out = (p @ shp(v)).permute(0, 2, 1, 3).reshape(t, cq).half()
        self.launches += 1
        return out
