"""Chained halo launches (`i2r_conv_halo_chain`, include/i2r.h): several dependent layers in ONE persistent grid, ordered
by per-image completion counters instead of launch boundaries.  The arithmetic of a tile is the same code either way, so
a chained run must be BIT-IDENTICAL to the same layers launched one by one -- any ordering hole (a tile reading an image
before its producer stored it) shows up as a changed result; every case is repeated to catch timing-dependent ones.
Shapes follow the reference: BasicBlock sequences of an HRNet module (lib/models/interformer_pureMulti.py:37-66,
:284-330), layer1's Bottlenecks (:69-107).
"""
import types

import pytest
import torch

import paths  # noqa: F401
from i2r_b200 import capi
from i2r_b200.hrnet_w48 import BackboneProgram
from i2r_b200.ops import ConvLayer, Runner, split_precision
from i2r_b200.packing import conv_taps, split_pair

pytestmark = pytest.mark.gpu


def _layer(g, cin, cout, k, relu, dev, split):
    w = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * (3.0 / (k * k * cin)) ** 0.5
    mats, dys, dxs = conv_taps(w, pad=k // 2)
    with split_precision(split):
        return ConvLayer(mats, dys, dxs, torch.rand(cout, generator=g) * 0.5 + 0.75, torch.rand(cout, generator=g) - 0.5,
                         relu=relu, device=dev)


def _act(g, shape, dev, split):
    x = torch.randn(*shape, generator=g)
    return (split_pair(x) if split else x.half()).to(dev)


def _module(g, chans, depth, dev, split):
    units = [[types.SimpleNamespace(c1=_layer(g, c, c, 3, True, dev, split), c2=_layer(g, c, c, 3, True, dev, split))
              for _ in range(depth)] for c in chans]
    return types.SimpleNamespace(units=units, nb=len(chans))


def _run(r, fn, chained):
    r.chain_enabled = chained
    before = r.chains
    with r.chain():
        out = fn()
    torch.cuda.synchronize()
    return out, r.chains - before


@pytest.mark.parametrize("split", [False, True], ids=["fp16", "split"])
@pytest.mark.parametrize("crops", [3, 32])
def test_basicblock_module_chain_is_bit_identical(crops, split):
    """Three resolution branches x 4 BasicBlocks = 8 dependent layers of 3-5 problems (192-channel layers run as two
    half-width problems writing channel slices of one tensor: two producers per consumer)."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(100 + crops)
    chans = (48, 96, 192)
    mod = _module(g, chans, 4, dev, split)
    xs = [_act(g, (crops, 64 >> b, 48 >> b, c), dev, split) for b, c in enumerate(chans)]
    r = Runner(dev, 0)
    r.split = split
    ref, n0 = _run(r, lambda: BackboneProgram._branches(r, mod, list(xs)), False)
    assert n0 == 0
    for it in range(30):
        out, n1 = _run(r, lambda: BackboneProgram._branches(r, mod, list(xs)), True)
        assert n1 >= 1, "the module must run as a chained launch"
        for a, b in zip(out, ref):
            assert torch.equal(a, b), "iteration %d: chained result differs" % it
    for o in out:
        assert torch.isfinite(o.float()).all()


@pytest.mark.parametrize("hw", [(64, 48), (16, 12), (20, 12)], ids=["64x48", "16x12_strips_straddle_images", "20x12_ragged"])
def test_bottleneck_chain_is_bit_identical(hw):
    """1x1 (pixel strip tiling) -> 3x3 (image tiling) -> 1x1 + residual, four Bottlenecks, the first with a downsample
    branch: strip tiles that straddle two images (16x12 = 192 pixels = 1.5 tiles) and image tiles clipped at the border."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(hw[0])
    units = []
    cin = 64
    for i in range(4):
        units.append(types.SimpleNamespace(
            c1=_layer(g, cin, 64, 1, True, dev, False), c2=_layer(g, 64, 64, 3, True, dev, False),
            c3=_layer(g, 64, 256, 1, True, dev, False), ds=_layer(g, cin, 256, 1, False, dev, False) if i == 0 else None))
        cin = 256
    x = _act(g, (12, hw[0], hw[1], 64), dev, False)
    r = Runner(dev, 0)

    def fwd():
        h = x
        for u in units:
            h = BackboneProgram._bottleneck(r, u, h)
        return h
    ref, _ = _run(r, fwd, False)
    for it in range(30):
        out, n = _run(r, fwd, True)
        assert n == 1
        assert torch.equal(out, ref), "iteration %d: chained result differs" % it
    assert torch.isfinite(out.float()).all()


def test_chain_falls_back_when_outputs_alias():
    """A layer writing into a tensor an earlier layer of the chain reads is not ordered by the counters: the entry point
    refuses (I2R_E_UNSUPPORTED) and the runner launches layer by layer -- same result."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    L1, L2 = _layer(g, 48, 48, 3, True, dev, False), _layer(g, 48, 48, 3, True, dev, False)
    x = _act(g, (4, 64, 48, 48), dev, False)
    r = Runner(dev, 0)
    h = r.conv(L1, x)
    ref = r.conv(L2, h)
    torch.cuda.synchronize()
    x2 = x.clone()
    r.chain_enabled = True
    before = r.chains
    with r.chain():
        h2 = r.conv(L1, x2)
        y = r.conv(L2, h2, out=x2)         # overwrites the input of the first layer
    torch.cuda.synchronize()
    assert r.chains == before
    assert torch.equal(y, ref)


def test_chain_workspace_and_limits():
    lib = capi.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(6)
    L = _layer(g, 48, 48, 3, True, dev, False)
    x = _act(g, (5, 64, 48, 48), dev, False)
    r = Runner(dev, 0)
    p, _ = r.problem(L, x)
    arr = (capi.ConvProblem * 2)(p, p)
    assert lib.i2r_conv_halo_chain_workspace(arr, 2) == 2 * 5 * 4
    import ctypes
    counts = (ctypes.c_int * 1)(2)
    ws = torch.zeros(16, dtype=torch.int32, device=dev)
    assert lib.i2r_conv_halo_chain(arr, counts, 1, ws.data_ptr(), 8, None) == -1       # workspace too small
    assert lib.i2r_conv_halo_chain(arr, counts, 0, ws.data_ptr(), 64, None) == -1      # no layers


def test_c2_forward_chained_equals_unchained():
    """Whole vanilla model (C2 shape, 8 crops): heatmaps with chained launches == heatmaps without, bit for bit, and the
    chained forward makes fewer launches."""
    from helpers import build_model, inputs_for
    _, model, _ = build_model()
    model = model.cuda()
    model.use_cuda_graph = False
    model.check_impl = False
    model.prepare("cuda:0")
    length = [3, 1, 4]
    x, pm = inputs_for(length)
    r = model._program.runner
    r.chain_enabled = True              # (off by default, see ops.Runner)
    with torch.no_grad():
        a = model(x, pm, length).clone()
        assert r.chains > 0
        n_chained = r.launches
        r.chain_enabled = False
        r.launches = 0
        b = model(x, pm, length).clone()
        n_plain = r.launches
        r.chain_enabled = True
        for _ in range(10):
            assert torch.equal(model(x, pm, length), a)
    assert torch.equal(a, b)
    assert n_chained < n_plain
