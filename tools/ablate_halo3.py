"""In-kernel (clock64) timing of one CTA of the persistent conv kernel under the i2r_debug_flags ablations."""
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
NB = int(os.environ.get("NB", "128"))
cap = 1024
for (c, h, w) in ((48, 64, 48), (96, 32, 24), (192, 16, 12)):
    L, _, _, _ = t._mk_conv(c, c, 3, 1, True, dev, c)
    x = torch.randn(NB, h, w, c).to(dev).half()
    for flags in (0, 2, 1, 4, 5, 8, 9, 13):
        r.lib.i2r_debug_flags(flags)
        r.conv_group([(L, x, {})])
        torch.cuda.synchronize()
        buf = torch.zeros(4 * 2 * cap, dtype=torch.int64, device=dev)
        r.lib.i2r_debug_trace(ctypes.c_void_p(buf.data_ptr()), cap, 3)
        r.conv_group([(L, x, {})])
        torch.cuda.synchronize()
        r.lib.i2r_debug_trace(None, 0, 0)
        b = buf.cpu().tolist()
        ev = sorted((b[2 * i + 1], b[2 * i] >> 32, b[2 * i] & 0xffffffff) for i in range(4 * cap) if b[2 * i + 1])
        t0 = [e for e in ev if e[1] == 30][0][0]
        t1 = [e for e in ev if e[1] == 31][0][0]
        commits = [e[0] for e in ev if e[1] == 12]
        first = [e for e in ev if e[1] == 11][0][0]
        n = len(commits)
        period = (commits[-1] - commits[n // 4]) / max(1, (n - 1 - n // 4))
        nm = 9 * (c // 16)
        print("C=%3d %2dx%2d NB=%d dbg=%2d : kernel %7d clk, first operands at %5d, %3d tiles, steady %6.0f clk/tile = %5.1f clk/MMA" % (
            c, h, w, NB, flags, t1 - t0, first - t0, n, period, period / nm))
r.lib.i2r_debug_flags(0)
