"""Timing of one HRNet stage-3 module's BasicBlock layers (3 branches x 4 blocks = 8 dependent layers, 32 crops) as
separate launches and as one chained launch, with the chain ablations of i2r_debug_chain_flags."""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import paths  # noqa: E402,F401
from i2r_b200 import capi  # noqa: E402
from i2r_b200.hrnet_w48 import BackboneProgram  # noqa: E402
from i2r_b200.ops import ConvLayer, Runner  # noqa: E402
from i2r_b200.packing import conv_taps  # noqa: E402

crops = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)


def layer(c):
    w = (torch.rand(c, c, 3, 3, generator=g) * 2 - 1) * (3.0 / (9 * c)) ** 0.5
    mats, dys, dxs = conv_taps(w, pad=1)
    return ConvLayer(mats, dys, dxs, torch.ones(c), torch.zeros(c), relu=True, device=dev)


chans = (48, 96, 192)
mod = types.SimpleNamespace(nb=3, units=[[types.SimpleNamespace(c1=layer(c), c2=layer(c)) for _ in range(4)] for c in chans])
xs = [torch.randn(crops, 64 >> b, 48 >> b, c, generator=g).half().to(dev) for b, c in enumerate(chans)]
r = Runner(dev, 0)
lib = capi.load()


def run(chained, iters=20, depth=4):
    r.chain_enabled = chained
    m = types.SimpleNamespace(nb=3, units=[u[:depth] for u in mod.units])

    def once():
        with r.chain():
            BackboneProgram._branches(r, m, list(xs))
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(graph, stream=st):
            for _ in range(iters):
                once()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


for depth in (1, 2, 4):
    print("depth %d blocks: separate %7.1f us, chained %7.1f us" % (depth, run(False, depth=depth), run(True, depth=depth)))
print("separate launches : %7.1f us per module" % run(False))
for flags, what in ((0, "product"), (1, "no consumer waits"), (2, "polls without fences"), (4, "publish after .read"),
                    (5, "no waits + publish after .read"), (9, "no waits, no publish"), (13, "no waits, no publish, .read"),
                    (17, "no waits, relaxed publish without fences"), (21, "no waits, relaxed publish, .read")):
    lib.i2r_debug_chain_flags(flags)
    print("chained, flags %d   : %7.1f us per module  (%s)" % (flags, run(True), what))
lib.i2r_debug_chain_flags(0)
