"""Per-CTA trace summary of ONE conv_halo launch of an arbitrary layer (Runner.conv: column chunking / N-split as in the
model).  usage: trace_halo_problem.py cin cout taps split gelu residual crops h w [step]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import paths  # noqa: E402,F401
from i2r_b200.ops import ConvLayer, Runner, split_precision  # noqa: E402
from i2r_b200.packing import split_pair  # noqa: E402

cin, cout, taps, split, gelu, residual, crops, h, w = (int(a) for a in sys.argv[1:10])
step = int(sys.argv[10]) if len(sys.argv) > 10 else 4
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
if taps == 9:
    mats = [(torch.rand(cout, cin, generator=g) * 2 - 1) / (9 * cin) ** 0.5 for _ in range(9)]
    dys, dxs = [t // 3 - 1 for t in range(9)], [t % 3 - 1 for t in range(9)]
else:
    mats, dys, dxs = [(torch.rand(cout, cin, generator=g) * 2 - 1) / cin ** 0.5], [0], [0]
with split_precision(bool(split)):
    L = ConvLayer(mats, dys, dxs, torch.ones(cout), torch.zeros(cout), relu=not gelu, device=dev)
r = Runner(dev, 0)
x32 = torch.randn(crops, h, w, cin, generator=g)
a32 = torch.randn(crops, h, w, cout, generator=g)
x = (split_pair(x32) if split else x32.half()).to(dev)
a = ((split_pair(a32) if split else a32.half()).to(dev)) if residual else None
kw = dict(add0=a, gelu=bool(gelu), act_first=bool(gelu and residual))
for _ in range(3):
    r.conv(L, x, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    r.conv(L, x, **kw)
e1.record()
torch.cuda.synchronize()
print("avg launch %.1f us (10 back to back)" % (e0.elapsed_time(e1) * 100))
cap = 2048
rows = []
for cta in range(0, 148, step):
    buf = torch.zeros(4 * 2 * cap, dtype=torch.int64, device=dev)
    r.lib.i2r_debug_trace(ctypes.c_void_p(buf.data_ptr()), cap, cta)
    r.conv(L, x, **kw)
    torch.cuda.synchronize()
    r.lib.i2r_debug_trace(None, 0, 0)
    b = buf.cpu().tolist()
    ev = sorted((b[2 * i + 1], b[2 * i] >> 32, b[2 * i] & 0xffffffff) for i in range(4 * cap) if b[2 * i + 1])
    if not ev:
        continue
    t0 = ev[0][0]
    commits = [c - t0 for c, tag, _ in ev if tag == 12]
    stored = [c - t0 for c, tag, _ in ev if tag == 21]
    landed = [c - t0 for c, tag, _ in ev if tag == 11]
    end = [c - t0 for c, tag, _ in ev if tag == 31]
    rows.append((cta, len(stored), landed[0] if landed else -1, commits[0] if commits else -1, stored[0] if stored else -1,
                 (stored[-1] - stored[0]) / max(1, len(stored) - 1) if len(stored) > 1 else 0, end[-1] if end else -1))
print("cta  tiles  first_A_landed  first_commit  first_stored  cycles_per_tile  span")
for row in rows:
    print("%3d  %5d  %14d  %12d  %12d  %15.0f  %6d" % row)
