"""Hardware bring-up probe: 3x3 conv on the persistent kernel vs the scalar check kernel with the UMMA
descriptor base_offset for dx-shifted windows switched on (default) and off (I2R_DESC_BASE_OFFSET=0)."""
import os
import subprocess
import sys

CODE = r"""
import sys, math
sys.path.insert(0, %r)
import paths, torch
sys.path.insert(0, paths.REPO + '/tests')
import test_kernels_gpu as t
from i2r_b200.ops import Runner
dev = torch.device('cuda:0')
r, c = Runner(dev, 0), Runner(dev, 1)
for (cin, cout, k) in ((48, 48, 3), (64, 64, 3), (96, 96, 3), (192, 96, 1), (192, 192, 3)):
    L, w, s, b = t._mk_conv(cout, cin, k, 1, True, dev, 1)
    x = torch.randn(4, 32, 24, cin).to(dev).half()
    d = (r.conv(L, x).float() - c.conv(L, x).float()).abs().max().item()
    torch.cuda.synchronize()
    print('cin', cin, 'cout', cout, 'k', k, 'maxdiff', d)
"""
repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for mode in ("1", "0"):
    env = dict(os.environ, I2R_DESC_BASE_OFFSET=mode)
    out = subprocess.run([sys.executable, "-c", CODE % repo], env=env, capture_output=True, text=True, timeout=300)
    print("=== I2R_DESC_BASE_OFFSET=%s rc=%d" % (mode, out.returncode))
    print(out.stdout[-1500:], out.stderr[-800:])
