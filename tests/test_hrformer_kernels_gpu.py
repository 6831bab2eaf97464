"""GPU parity of the HRFormer-B building-block kernels (depthwise 3x3 + BN + GELU, bilinear fuse, padded LayerNorm)
against float64 torch restatements of the reference ops (lib/models/hrformer.py:1094-1119, :1616-1731, :1198)."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

import paths

pytestmark = pytest.mark.gpu
REPORT = os.path.join(paths.REPO, "gpurun_out", "kernel_report.jsonl")


def _report(**kw):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(json.dumps(kw) + "\n")


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def tc(dev):
    from i2r_b200.ops import Runner
    return Runner(dev, impl=0)


def _enc(v, split):
    from i2r_b200.packing import split_pair
    return split_pair(v) if split else v.half()


def _dec(v, split):
    from i2r_b200.packing import merge_pair
    return merge_pair(v.cpu()).double() if split else v.cpu().double()


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("c,h,w,stride,act", [(312, 64, 48, 1, "gelu"), (80, 32, 24, 2, None), (640, 7, 5, 1, "relu"),
                                              (160, 96, 72, 2, "gelu")])
def test_dwconv3x3(dev, tc, split, c, h, w, stride, act):
    g = torch.Generator().manual_seed(c + h + stride)
    nb = 2
    x32 = torch.randn(nb, h, w, c, generator=g)
    wt = (torch.rand(c, 1, 3, 3, generator=g) * 2 - 1) / 3
    sc = torch.rand(c, generator=g) + 0.5
    bi = torch.randn(c, generator=g) * 0.1
    x = _enc(x32, split)
    tc.split = split
    try:
        y = tc.dwconv3x3(x.to(dev), wt.reshape(c, 9).t().contiguous().to(dev), sc.to(dev), bi.to(dev), stride, act)
        torch.cuda.synchronize()
    finally:
        tc.split = False
    xin = _dec(x, split).permute(0, 3, 1, 2)
    ref = F.conv2d(xin, wt.double(), None, stride, 1, 1, c) * sc.double().view(1, -1, 1, 1) + bi.double().view(1, -1, 1, 1)
    ref = {"gelu": F.gelu, "relu": F.relu, None: lambda t: t}[act](ref).permute(0, 2, 3, 1)
    err = float((_dec(y, split) - ref).abs().max())
    _report(test="dwconv3x3", split=split, c=c, stride=stride, act=act, err=err)
    assert tuple(y.shape) == (nb, (h + stride - 1) // stride, (w + stride - 1) // stride, (2 if split else 1) * c)
    assert err <= (5e-6 if split else 4e-3), err


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("shifts", [(1,), (1, 2), (1, 2, 3)])
def test_upsum_bilinear(dev, tc, split, shifts):
    g = torch.Generator().manual_seed(10 * len(shifts))
    nb, h, w, c = 2, 64, 48, 80
    x32 = torch.randn(nb, h, w, c, generator=g)
    ts32 = [torch.randn(nb, h >> s, w >> s, c, generator=g) for s in shifts]
    x0, ts = _enc(x32, split), [_enc(t, split) for t in ts32]
    tc.split = split
    try:
        y = tc.upsum_bilinear(x0.to(dev), [(t.to(dev), s) for t, s in zip(ts, shifts)], relu=True)
        torch.cuda.synchronize()
    finally:
        tc.split = False
    ref = _dec(x0, split).permute(0, 3, 1, 2)
    for t, s in zip(ts, shifts):
        ref = ref + F.interpolate(_dec(t, split).permute(0, 3, 1, 2), scale_factor=2 ** s, mode="bilinear",
                                  align_corners=False)
    ref = F.relu(ref).permute(0, 2, 3, 1)
    err = float((_dec(y, split) - ref).abs().max())
    _report(test="upsum_bilinear", split=split, shifts=list(shifts), err=err)
    assert err <= (5e-6 if split else 4e-3), err


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("c_real,c_pad", [(78, 80), (156, 160), (312, 320), (624, 640), (96, 96)])
def test_layernorm_padded(dev, tc, split, c_real, c_pad):
    g = torch.Generator().manual_seed(c_real)
    rows = 1001
    x32 = torch.randn(rows, c_pad, generator=g) * 2 + 0.3
    x32[:, c_real:] = 7.0            # garbage in the pad channels must not leak into the statistics
    gamma = torch.rand(c_pad, generator=g) + 0.5
    beta = torch.randn(c_pad, generator=g) * 0.1
    x = _enc(x32, split)
    tc.split = split
    try:
        y = tc.layernorm_padded(x.to(dev), gamma.to(dev), beta.to(dev), c_real, eps=1e-6)
        torch.cuda.synchronize()
    finally:
        tc.split = False
    xin = _dec(x, split)[:, :c_real]
    ref = F.layer_norm(xin, (c_real,), gamma[:c_real].double(), beta[:c_real].double(), 1e-6)
    got = _dec(y, split)
    err = float((got[:, :c_real] - ref).abs().max())
    _report(test="layernorm_padded", split=split, c_real=c_real, err=err)
    assert err <= (5e-6 if split else 4e-3), err
    assert float(got[:, c_real:].abs().max()) == 0.0 if c_pad > c_real else True


def _window_rows_ref(x, ws=7):
    """[NB, H, W, C] -> window-major rows [NB*QH*QW*ws*ws, C] with centre zero padding (hrformer.py:949-986)."""
    nb, h, w, c = x.shape
    ph, pw = (-h) % ws, (-w) % ws
    xp = F.pad(x, (0, 0, pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
    hp, wp = h + ph, w + pw
    return xp.view(nb, hp // ws, ws, wp // ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(-1, c), (ph, pw, hp, wp)


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("c_real,c_pad,h,w", [(78, 80, 64, 48), (156, 160, 32, 24), (624, 640, 8, 6), (78, 80, 96, 72)])
def test_ln_window_gather_and_scatter(dev, tc, split, c_real, c_pad, h, w):
    g = torch.Generator().manual_seed(c_real + h)
    nb = 2
    x32 = torch.randn(nb, h, w, c_pad, generator=g) * 1.5 + 0.2
    x32[..., c_real:] = 0
    gamma = torch.rand(c_pad, generator=g) + 0.5
    beta = torch.randn(c_pad, generator=g) * 0.1
    x = _enc(x32, split)
    tc.split = split
    try:
        rows = tc.ln_window_gather(x.to(dev), gamma.to(dev), beta.to(dev), c_real)
        back = tc.window_scatter_add(x.to(dev), rows)
        torch.cuda.synchronize()
    finally:
        tc.split = False
    xin = _dec(x, split)
    ln = torch.zeros_like(xin)
    ln[..., :c_real] = F.layer_norm(xin[..., :c_real], (c_real,), gamma[:c_real].double(), beta[:c_real].double(), 1e-6)
    ref_rows, _ = _window_rows_ref(ln)
    got = _dec(rows, split)
    err = float((got - ref_rows).abs().max())
    # scatter-add: x + LN(x) at every pixel (the window rows hold LN(x); padded rows are dropped)
    err2 = float((_dec(back, split) - (xin + ln)).abs().max())
    _report(test="ln_window_gather", split=split, c=c_real, h=h, err=err, err_scatter=err2)
    assert tuple(got.shape) == tuple(ref_rows.shape)
    assert err <= (5e-6 if split else 4e-3) and err2 <= (1e-5 if split else 8e-3), (err, err2)


@pytest.mark.parametrize("kernel", ["tcgen05", "mma_sync"])
@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("heads,nwin", [(2, 140), (16, 4), (4, 7), (8, 1)])
def test_window_attention(dev, tc, split, heads, nwin, kernel):
    """49-token windows, head_dim 39 padded to 48 (pad channels zero), scale 39**-0.5, no bias / mask.  Product kernel
    (tcgen05: two windows per 128-row tile, odd window counts leave half a tile empty) and the mma.sync check kernel."""
    from i2r_b200.packing import merge_pair, split_pair
    g = torch.Generator().manual_seed(heads)
    hd, hp, wl = 39, 48, 49
    t = nwin * wl
    qkv32 = torch.randn(3, t, heads, hp, generator=g)
    qkv32[..., hd:] = 0
    qkv32 = qkv32.reshape(3, t, heads * hp)
    scale = hd ** -0.5
    if split:
        q, k, v = (split_pair(m) for m in qkv32)
        vals = [merge_pair(m).double() for m in (q, k, v)]
        cq = heads * hp
        args = [m.to(dev)[:, :cq] for m in (q, k, v)]
    else:
        q, k, v = (m.half() for m in qkv32)
        vals = [m.double() for m in (q, k, v)]
        args = [m.to(dev) for m in (q, k, v)]
    tc.split = split
    tc.window_att_tc = kernel == "tcgen05"
    try:
        out = tc.window_attention(args[0], args[1], args[2], wl, heads, scale)
        torch.cuda.synchronize()
        again = tc.window_attention(args[0], args[1], args[2], wl, heads, scale)
        torch.cuda.synchronize()
        assert torch.equal(out, again)
    finally:
        tc.split = False
        tc.window_att_tc = True
    qd, kd, vd = (m.view(nwin, wl, heads, hp).permute(0, 2, 1, 3) for m in vals)
    ref = (torch.softmax(qd @ kd.transpose(-1, -2) * scale, dim=-1) @ vd).permute(0, 2, 1, 3).reshape(t, heads * hp)
    got = merge_pair(out.cpu()).double() if split else out.cpu().double()
    err = float((got - ref).abs().max())
    _report(test="window_attention", split=split, heads=heads, nwin=nwin, kernel=kernel, err=err)
    assert err <= (4e-4 if split else 3e-3), err


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("path", ["halo", "igemm", "check"])
@pytest.mark.parametrize("act_first", [False, True])
def test_gemm_epilogue_gelu_and_act_first(dev, split, path, act_first):
    """1x1 conv + folded BN + erf-GELU with a residual added before or after the activation (fc1 / fc2 of MlpDWBN)."""
    import math
    from i2r_b200.ops import ConvLayer, Runner, split_precision
    g = torch.Generator().manual_seed(3 + int(act_first))
    nb, h, w, cin, cout = 2, 16, 8, 80, 96
    wt = (torch.rand(cout, cin, generator=g) * 2 - 1) / math.sqrt(cin)
    sc = torch.rand(cout, generator=g) + 0.5
    bi = torch.randn(cout, generator=g) * 0.1
    x32 = torch.randn(nb, h, w, cin, generator=g)
    a32 = torch.randn(nb, h, w, cout, generator=g)
    r = Runner(dev, impl=1 if path == "check" else 0)
    r.use_tma = path == "halo"
    with split_precision(split):
        L = ConvLayer([wt], [0], [0], sc, bi, device=dev)
    x, a = _enc(x32, split).to(dev), _enc(a32, split).to(dev)
    p, out = r.problem(L, x, add0=a, relu=False, gelu=True, act_first=act_first)
    r.launch([p])
    torch.cuda.synchronize()
    lin = (_dec(x, split) @ wt.double().t()) * sc.double() + bi.double()
    ref = F.gelu(lin) + _dec(a, split) if act_first else F.gelu(lin + _dec(a, split))
    err = float((_dec(out, split) - ref).abs().max())
    _report(test="gemm_gelu", split=split, path=path, act_first=act_first, err=err)
    assert err <= (3e-5 if split else 6e-3), err
