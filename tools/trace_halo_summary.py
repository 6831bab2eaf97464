"""Per-CTA summary of the persistent conv kernel on a stage-3 BasicBlock group (three resolution branches, batch 32,
N-split routing as in the model): for every CTA the number of tiles it ran, the cycle of its first MMA commit and its
K.start -> K.end span.  Shows start-up cost and load imbalance of the static CTA allocation."""
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
specs = []
for (c, h, w) in ((48, 64, 48), (96, 32, 24), (192, 16, 12)):
    L, _, _, _ = t._mk_conv(c, c, 3, 1, True, dev, c)
    x = torch.randn(32, h, w, c).to(dev).half()
    specs.append((L, x, {"add0": torch.randn(32, h, w, c).to(dev).half()} if "--res" in sys.argv else {}))
for _ in range(3):
    r.conv_group(specs)
torch.cuda.synchronize()
cap = 1024
rows = []
for cta in range(0, 148, int(os.environ.get("STEP", "3"))):
    buf = torch.zeros(4 * 2 * cap, dtype=torch.int64, device=dev)
    r.lib.i2r_debug_trace(ctypes.c_void_p(buf.data_ptr()), cap, cta)
    r.conv_group(specs)
    torch.cuda.synchronize()
    r.lib.i2r_debug_trace(None, 0, 0)
    b = buf.cpu().tolist()
    ev = sorted((b[2 * i + 1], b[2 * i] >> 32, b[2 * i] & 0xffffffff) for i in range(4 * cap) if b[2 * i + 1])
    if not ev:
        continue
    t0 = ev[0][0]
    commits = [c - t0 for c, tag, _ in ev if tag == 12]
    stored = [c - t0 for c, tag, _ in ev if tag == 21]
    end = [c - t0 for c, tag, _ in ev if tag == 31]
    rows.append((cta, len(stored), commits[0] if commits else -1, stored[0] if stored else -1,
                 (stored[-1] - stored[0]) / max(1, len(stored) - 1) if len(stored) > 1 else 0, end[-1] if end else -1))
print("cta  tiles  first_commit  first_stored  cycles_per_tile  span")
for row in rows:
    print("%3d  %5d  %12d  %12d  %15.0f  %6d" % row)
