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
from i2r_b200.packing import unpack_taps


class EmuRunner(Runner):
    def __init__(self):
        self.impl = -1
        self.use_tma = True     # exercise the N-split logic of Runner.problems
        self.timing = None
        self.launches = 0
        self.device = torch.device("cpu")

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
        w = unpack_taps(L.w, cin)                         # [ntaps, npad, cin]
        oy = torch.arange(p.OH).view(-1, 1).expand(p.OH, p.OW)
        ox = torch.arange(p.OW).view(1, -1).expand(p.OH, p.OW)
        acc = torch.zeros(nb, p.OH, p.OW, npad)
        xf = x[..., :cin].float()
        for t in range(p.ntaps):
            wt = w[t]
            iy = oy * p.stride + int(p.dy[t])
            ix = ox * p.stride + int(p.dx[t])
            ok = (iy >= 0) & (iy < p.IH) & (ix >= 0) & (ix < p.IW)
            sy = (iy.clamp(0, p.IH - 1) >> sh)
            sx = (ix.clamp(0, p.IW - 1) >> sh)
            a = xf[:, sy, sx, :] * ok[None, :, :, None]
            acc += a @ wt.t()
        v = acc[..., :cout] * L.scale[:cout] + L.bias[:cout]
        fy = oy * p.out_mul + p.out_offy
        fx = ox * p.out_mul + p.out_offx
        for a, s in ((add0, p.add0_shift), (add1, p.add1_shift)):
            if a is not None:
                assert tuple(a.shape) == (nb, p.OHf >> s, p.OWf >> s, cout) and a.stride(2) == p.add_pix_stride
                v = v + a.float()[:, fy >> s, fx >> s, :]
        if p.flags & capi.F_RELU:
            v = F.relu(v)
        if p.flags & capi.F_OUT_NCHW_F32:
            out[:, :, fy, fx] = v.permute(0, 3, 1, 2)
        elif p.flags & capi.F_OUT_F32:
            out[:, fy, fx, :] = v
        else:
            out[:, fy, fx, :] = v.half()

    def stem(self, x, w, scale, bias, cout):
        cin = x.shape[1]
        wt = w.reshape(cin, 3, 3, cout).permute(3, 0, 1, 2)
        y = F.conv2d(x, wt, None, 2, 1) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
        self.launches += 1
        return F.relu(y).permute(0, 2, 3, 1).contiguous().half()

    def maxpool(self, x):
        self.launches += 1
        return F.max_pool2d(x.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).contiguous().half()

    def layernorm(self, x2d, gamma, beta, eps=1e-5, pos=None):
        y = F.layer_norm(x2d.float(), (x2d.shape[1],), gamma, beta, eps)
        self.launches += 1
        return y.half(), ((y + pos.float()).half() if pos is not None else None)

    def add(self, a, b):
        self.launches += 1
        return (a.float() + b.float()).half()

    def attention(self, q, k, v, cu_seqlens, max_seqlen, scale):
        out = torch.empty(q.shape[0], q.shape[1], dtype=torch.float16)
        cu = cu_seqlens.tolist()
        for a, b in zip(cu[:-1], cu[1:]):
            assert b - a <= max_seqlen
            s = torch.softmax(q[a:b].float() @ k[a:b].float().t() * scale, dim=-1)
            out[a:b] = (s.half().float() @ v[a:b].float()).half()
        self.launches += 1
        return out
