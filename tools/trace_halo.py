"""Event timeline of one CTA of the persistent conv kernel on a stage-3-like grouped launch
(three resolution branches, batch 32).  Prints per-tile timestamps (SM cycles) for each role."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import paths  # noqa: E402,F401

sys.path.insert(0, os.path.join(paths.REPO, "tests"))
import test_kernels_gpu as t  # noqa: E402
from i2r_b200.ops import Runner  # noqa: E402

dev = torch.device("cuda:0")
r = Runner(dev, 0)
r.lib.i2r_debug_flags(int(os.environ.get('DBG', '0')))
specs = []
for (c, h, w) in ((48, 64, 48), (96, 32, 24), (192, 16, 12))[:int(os.environ.get('NPROB', '3'))]:
    L, _, _, _ = t._mk_conv(c, c, 3, 1, True, dev, c)
    x = torch.randn(32, h, w, c).to(dev).half()
    specs.append((L, x, {"add0": torch.randn(32, h, w, c).to(dev).half()} if "--res" in sys.argv else {}))
for _ in range(3):
    r.conv_group(specs)
torch.cuda.synchronize()
for cta in [int(a) for a in sys.argv[1:] if a.isdigit()] or [0]:
    cap = 1024
    buf = torch.zeros(4 * 2 * cap, dtype=torch.int64, device=dev)
    r.lib.i2r_debug_trace(ctypes.c_void_p(buf.data_ptr()), cap, cta)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r.conv_group(specs)
    e1.record()
    torch.cuda.synchronize()
    r.lib.i2r_debug_trace(None, 0, 0)
    b = buf.cpu().tolist()
    ev = sorted((b[2 * i + 1], b[2 * i] >> 32, b[2 * i] & 0xffffffff) for i in range(4 * cap) if b[2 * i + 1])
    n = len(ev)
    t0 = ev[0][0] if ev else 0
    print("=== cta %d: %d events, launch %.1f us" % (cta, n, e0.elapsed_time(e1) * 1e3))
    names = {1: "P.free", 2: "P.issued", 3: "P.publish", 10: "M.accfree", 11: "M.operands", 12: "M.commit",
             20: "E.ready", 21: "E.stored", 27: "E.top", 28: "E.waited", 22: "E.loaded", 24: "E.data", 25: "E.packed0", 26: "E.stg0", 23: "E.math", 30: "K.start", 31: "K.end", 13: "M.weights"}
    for clk, tag, tile in ev[:int(os.environ.get("TRACE_MAX", "120"))]:
        print("%8d  %-11s tile %d" % (clk - t0, names.get(tag, tag), tile))
