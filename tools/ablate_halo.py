"""Times single-branch stride-1 3x3 launches of the persistent conv kernel (batch 32) under the
i2r_debug_flags ablations: which role paces the tile period?"""
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
for (c, h, w) in ((48, 64, 48), (96, 32, 24), (192, 16, 12), (64, 64, 48)):
    L, _, _, _ = t._mk_conv(c, c, 3, 1, True, dev, c)
    x = torch.randn(32, h, w, c).to(dev).half()
    res = torch.randn(32, h, w, c).to(dev).half()
    for name, kw in (("plain", {}), ("residual", {"add0": res})):
        for flags in (0, 2, 1):
            r.lib.i2r_debug_flags(flags)
            for _ in range(3):
                r.conv_group([(L, x, kw)])
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(20):
                    r.conv_group([(L, x, kw)])
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 20
            fl = 2.0 * 32 * h * w * c * c * 9
            print("C=%3d %2dx%2d %-8s dbg=%d : %7.1f us  %6.1f TFLOP/s" % (c, h, w, name, flags, us, fl / us * 1e-6))
r.lib.i2r_debug_flags(0)
