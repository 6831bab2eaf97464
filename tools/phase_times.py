"""Per-launch GPU time of an eager forward queued behind a spin kernel (host launch gaps excluded), grouped by phase.
Events bracket every Runner entry point; PDL overlap between neighbouring kernels is attributed to the later one."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import paths  # noqa: E402,F401

sys.path.insert(0, os.path.join(paths.REPO, "tests"))
from helpers import build_model, inputs_for  # noqa: E402
from i2r_b200 import ops  # noqa: E402

yaml = sys.argv[1] if len(sys.argv) > 1 else "coco/interformer_coco_w48_pure_en6.yaml"
images, persons = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (8, 4)
cfg, model, sd = build_model(yaml)
model = model.cuda()
model.use_cuda_graph = False
length = [persons] * images
hw = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (256, 192)
x, pm = inputs_for(length, *hw)
x, pm = x.cuda(), pm.cuda()
model(x, pm, length)
torch.cuda.synchronize()

records = []
pending = [[]]


def wrap(name):
    orig = getattr(ops.Runner, name)

    def f(self, *a, **kw):
        if name == "_flush_chain":
            pending[0] = [list(layer) for layer in (self._chain or [])] if len(self._chain or []) > 1 else []
            if not pending[0]:       # nothing or a single layer: the inner _launch_now records it
                return orig(self, *a, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(self, *a, **kw)
        e1.record()
        desc = name
        if name == "_launch_now":
            desc = "conv[" + ",".join("%dx%d:%d>%d%s" % (p.IH, p.IW, p.Cin, p.Cout, "k%d" % p.ntaps) for p in a[0]) + "]"
        if name == "_flush_chain":
            if not pending[0]:
                return out
            desc = "chain[%d layers: %s]" % (len(pending[0]), " | ".join(
                ",".join("%dx%d:%d>%d%s" % (p.IH, p.IW, p.Cin, p.Cout, "k%d" % p.ntaps) for p in layer) for layer in pending[0][:2]))
        records.append((desc, e0, e1))
        return out
    setattr(ops.Runner, name, f)


for n in ("_launch_now", "_flush_chain", "stem", "maxpool", "layernorm", "add", "upsum", "attention", "attention_tc", "encoder_tail", "dwconv3x3",
          "upsum_bilinear", "layernorm_padded", "ln_window_gather", "window_scatter_add", "window_attention"):
    wrap(n)
reps = 3
tot = {}
for rep in range(reps):
    records.clear()
    torch.cuda._sleep(40_000_000)
    model(x, pm, length)
    torch.cuda.synchronize()
    for i, (d, a, b) in enumerate(records):
        tot.setdefault(i, [d, 0.0])[1] += a.elapsed_time(b) * 1e3 / reps
total = sum(v[1] for v in tot.values())
print("eager spin-parked forward: %.1f us over %d launches (%s, %d crops)" % (total, len(tot), yaml, sum(length)))
for i in sorted(tot):
    print("%3d %8.1f us  %s" % (i, tot[i][1], tot[i][0][:150]))
by_kind = {}
for i in tot:
    k = tot[i][0].split("[")[0]
    if k == "conv":
        k = "conv " + ("igemm/halo 1x1" if "k1" in tot[i][0] and "k9" not in tot[i][0] else "3x3 / mixed")
    by_kind.setdefault(k, [0, 0.0])
    by_kind[k][0] += 1
    by_kind[k][1] += tot[i][1]
print("by kind:")
for k, (n, t) in sorted(by_kind.items(), key=lambda kv: -kv[1][1]):
    print("  %-28s %4d launches %9.1f us  %5.1f%%" % (k, n, t, 100 * t / total))
