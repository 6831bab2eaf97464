"""Data dependence of the persistent conv kernel's tile rate: same launch, different activation values."""
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
NB = int(os.environ.get("NB", "256"))
for (c, h, w) in ((48, 64, 48), (64, 64, 48), (96, 32, 24), (192, 16, 12)):
    L, _, _, _ = t._mk_conv(c, c, 3, 1, True, dev, c)
    for name in ("randn", "ones", "zeros"):
        x = {"randn": torch.randn, "ones": torch.ones, "zeros": torch.zeros}[name](NB, h, w, c).to(dev).half()
        for flags in (0, 1):
            r.lib.i2r_debug_flags(flags)
            for _ in range(2):
                r.conv_group([(L, x, {})])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r.conv_group([(L, x, {})])
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3
            fl = 2.0 * NB * h * w * c * c * 9
            tiles = NB * ((h + 15) // 16) * ((w + 7) // 8)
            per_cta = -(-tiles // 148)
            clk = us * 1.965e3 / per_cta
            nm = 9 * (c // 16)
            print("C=%3d %2dx%2d NB=%d %-6s dbg=%d : %8.1f us %6.1f TFLOP/s  %6.0f clk/tile  %5.1f clk/MMA" % (
                c, h, w, NB, name, flags, us, fl / us * 1e-6, clk, clk / nm))
r.lib.i2r_debug_flags(0)
