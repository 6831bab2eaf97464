"""GPU parity tests of the individual sm_100a kernels, all through the C ABI (i2r_b200.capi):
tcgen05 implicit GEMM vs (a) the scalar check kernel on the same packed operands and (b) a plain
torch fp32 reference of the same op on fp16-rounded operands; attention / LayerNorm / pooling / stem
vs torch fp32.  Tolerances are stated per test."""
import json
import math
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
def runners(dev):
    """(product runner: TMA kernel where eligible, else gather kernel; scalar check kernel)."""
    from i2r_b200.ops import Runner
    return Runner(dev, impl=0), Runner(dev, impl=1)


@pytest.fixture(scope="module")
def gather_runner(dev):
    """tcgen05 gather kernel for every problem (TMA routing off)."""
    from i2r_b200.ops import Runner
    r = Runner(dev, impl=0)
    r.use_tma = False
    return r


def _mk_conv(cout, cin, k, stride, relu, dev, seed):
    from i2r_b200.ops import ConvLayer
    from i2r_b200.packing import conv_taps
    g = torch.Generator().manual_seed(seed)
    w = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) / math.sqrt(cin * k * k)
    scale = torch.rand(cout, generator=g) + 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    mats, dys, dxs = conv_taps(w, pad=k // 2)
    return ConvLayer(mats, dys, dxs, scale, bias, stride=stride, relu=relu, device=dev), w, scale, bias


def _diff(a, b):
    d = (a.float() - b.float()).abs()
    return float(d.max()), float(d.mean()), int((d > 1e-2).sum())


CONV_CASES = [
    # name, NB, H, W, Cin, Cout, k, stride, relu
    ("c48_3x3_s1", 2, 64, 48, 48, 48, 3, 1, True),
    ("c96_3x3_s1", 2, 32, 24, 96, 96, 3, 1, True),
    ("c192_3x3_s1", 3, 16, 12, 192, 192, 3, 1, False),
    ("c256to96_3x3_s2", 2, 64, 48, 256, 96, 3, 2, True),
    ("c64to256_1x1", 1, 64, 48, 64, 256, 1, 1, False),
    ("c256to64_1x1", 1, 64, 48, 256, 64, 1, 1, True),
    ("c192to96_1x1", 5, 16, 12, 192, 96, 1, 1, False),
    ("c64_3x3_s2_big", 2, 128, 96, 64, 64, 3, 2, True),
    ("ragged_M", 1, 10, 7, 48, 48, 3, 1, True),
    ("c96_3x3_s1_b32", 32, 32, 24, 96, 96, 3, 1, True),        # N-split, resident weights, persistent loop
    ("c192_3x3_s1_b32", 32, 16, 12, 192, 192, 3, 1, True),     # streamed weights, partial x tiles
    ("c256to48_3x3_s1", 4, 64, 48, 256, 48, 3, 1, True),       # streamed weights, 4 K-chunks
    ("c48_3x3_s1_b32", 32, 64, 48, 48, 48, 3, 1, True),        # 768 tiles over 148 persistent CTAs
    ("c96to192_1x1_tokens", 1, 6144, 1, 96, 192, 1, 1, True),  # nn.Linear shape (re-tiled 8-wide)
    ("odd_map_3x3", 3, 21, 13, 64, 64, 3, 1, False),           # ragged tiles in x and y
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_igemm_conv(case, dev, runners, gather_runner):
    name, nb, h, w, cin, cout, k, stride, relu = case
    tc, chk = runners
    L, wt, scale, bias = _mk_conv(cout, cin, k, stride, relu, dev, seed=hash(name) % 1000)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(nb, h, w, cin, generator=g).to(dev).half()
    res = None
    y_tc = tc.conv(L, x)
    y_ck = chk.conv(L, x)
    y_ga = gather_runner.conv(L, x)
    torch.cuda.synchronize()
    # torch fp32 reference on the same fp16-rounded operands
    xr = x.float().permute(0, 3, 1, 2)
    wr = wt.to(dev).half().float()
    ref = F.conv2d(xr, wr, None, stride, k // 2) * scale.to(dev).view(1, -1, 1, 1) + bias.to(dev).view(1, -1, 1, 1)
    if relu:
        ref = F.relu(ref)
    ref = ref.permute(0, 2, 3, 1)
    e_ck = _diff(y_ck, ref)
    e_tc = _diff(y_tc, ref)
    e_x = _diff(y_tc, y_ck)
    e_ga = _diff(y_ga, ref)
    _report(test="igemm_conv", case=name, check_vs_torch=e_ck, tc_vs_torch=e_tc, tc_vs_check=e_x, gather_vs_torch=e_ga)
    assert e_ga[0] <= 4e-3, ("tcgen05 gather kernel vs torch", e_ga)
    # fp16 output rounding: |y| <~ 4 -> half-ulp 2e-3
    assert e_ck[0] <= 4e-3, ("check kernel vs torch", e_ck)
    assert e_tc[0] <= 4e-3, ("tcgen05 kernel vs torch", e_tc, "vs check", e_x)


def test_igemm_addends_upsample_group(dev, runners):
    """HRNet fuse semantics: in_shift gather, two addends with upsample shifts, grouped launch."""
    tc, chk = runners
    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(2, 64, 48, 48, generator=g).to(dev).half()
    x1 = torch.randn(2, 32, 24, 96, generator=g).to(dev).half()
    x2 = torch.randn(2, 16, 12, 192, generator=g).to(dev).half()
    t02 = torch.randn(2, 16, 12, 48, generator=g).to(dev).half()
    t12 = torch.randn(2, 16, 12, 96, generator=g).to(dev).half()
    L01, w01, s01, b01 = _mk_conv(48, 96, 1, 1, True, dev, 11)
    L10, w10, s10, b10 = _mk_conv(96, 48, 3, 2, True, dev, 12)
    outs = {}
    for tag, r in (("tc", tc), ("ck", chk)):
        outs[tag] = r.conv_group([
            (L01, x1, dict(in_shift=1, add0=x0, add1=t02, add1_shift=2, relu=True)),
            (L10, x0, dict(add0=x1, add1=t12, add1_shift=1, relu=True)),
        ])
    torch.cuda.synchronize()

    def nchw(t):
        return t.float().permute(0, 3, 1, 2)

    def affine(y, s, b):
        return y * s.to(dev).view(1, -1, 1, 1) + b.to(dev).view(1, -1, 1, 1)

    r0 = affine(F.conv2d(F.interpolate(nchw(x1), scale_factor=2, mode="nearest"), w01.to(dev).half().float()), s01, b01)
    r0 = F.relu(r0 + nchw(x0) + F.interpolate(nchw(t02), scale_factor=4, mode="nearest")).permute(0, 2, 3, 1)
    r1 = affine(F.conv2d(nchw(x0), w10.to(dev).half().float(), None, 2, 1), s10, b10)
    r1 = F.relu(r1 + nchw(x1) + F.interpolate(nchw(t12), scale_factor=2, mode="nearest")).permute(0, 2, 3, 1)
    for i, ref in enumerate((r0, r1)):
        e_ck, e_tc = _diff(outs["ck"][i], ref), _diff(outs["tc"][i], ref)
        _report(test="igemm_fuse", idx=i, check_vs_torch=e_ck, tc_vs_torch=e_tc)
        assert e_ck[0] <= 8e-3, e_ck
        assert e_tc[0] <= 8e-3, e_tc


def test_igemm_deconv_and_head(dev, runners):
    """ConvTranspose2d 4x4 s2 p1 + BN + ReLU as four phases; 1x1 head with fp32 NCHW output."""
    from i2r_b200.ops import ConvLayer
    from i2r_b200.packing import deconv4x4s2_phase_taps
    tc, chk = runners
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 16, 12, 96, generator=g).to(dev).half()
    w = (torch.rand(96, 96, 4, 4, generator=g) * 2 - 1) / math.sqrt(96 * 4)
    scale = torch.rand(96, generator=g) + 0.5
    bias = torch.randn(96, generator=g) * 0.1
    ref = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), w.to(dev).half().float(), None, 2, 1)
    ref = F.relu(ref * scale.to(dev).view(1, -1, 1, 1) + bias.to(dev).view(1, -1, 1, 1)).permute(0, 2, 3, 1)
    for tag, r in (("tc", tc), ("ck", chk)):
        out = torch.zeros(3, 32, 24, 96, dtype=torch.float16, device=dev)
        specs = []
        for py in (0, 1):
            for px in (0, 1):
                mats, dys, dxs = deconv4x4s2_phase_taps(w, py, px)
                L = ConvLayer(mats, dys, dxs, scale, bias, relu=True, device=dev)
                specs.append((L, x, dict(out=out, out_hw=(32, 24), out_mul=2, out_off=(py, px))))
        r.conv_group(specs)
        torch.cuda.synchronize()
        e = _diff(out, ref)
        _report(test="deconv", impl=tag, err=e)
        assert e[0] <= 4e-3, (tag, e)
    # head: 96 -> 17, bias, fp32 NCHW
    Lh, wh, sh, bh = _mk_conv(17, 96, 1, 1, False, dev, 21)
    xh = torch.randn(2, 64, 48, 96, generator=g).to(dev).half()
    refh = F.conv2d(xh.float().permute(0, 3, 1, 2), wh.to(dev).half().float()) * sh.to(dev).view(1, -1, 1, 1) + \
        bh.to(dev).view(1, -1, 1, 1)
    # the persistent halo kernel folds the scale into the fp16 weights (one rounding of w*scale) and feeds the bias
    # through an fp16 hi+lo pair: its "same operands" reference rounds w*scale instead of w
    wf = (wh * sh.view(-1, 1, 1, 1)).to(dev).half().float()
    refh_folded = F.conv2d(xh.float().permute(0, 3, 1, 2), wf) + bh.to(dev).view(1, -1, 1, 1)
    for tag, r, rf in (("tc", tc, refh_folded), ("ck", chk, refh)):
        yh = r.conv(Lh, xh, out_mode="nchw32")
        torch.cuda.synchronize()
        e = _diff(yh, rf)
        _report(test="head", impl=tag, err=e)
        assert yh.dtype == torch.float32 and tuple(yh.shape) == (2, 17, 64, 48)
        assert e[0] <= 1e-4, (tag, e)
        assert _diff(yh, refh)[0] <= 2e-3, (tag, "vs unfolded reference")


def test_linear_strided_and_relu(dev, runners):
    from i2r_b200.ops import ConvLayer
    tc, chk = runners
    g = torch.Generator().manual_seed(9)
    t = 777
    buf = torch.randn(t, 192, generator=g).to(dev).half()
    w = (torch.rand(192, 96, generator=g) * 2 - 1) / math.sqrt(96)
    b = torch.randn(192, generator=g) * 0.1
    L = ConvLayer([w], [0], [0], torch.ones(192), b, relu=True, device=dev)
    xin = buf[:, 96:]                      # strided view: row stride 192, 96 columns
    add = torch.randn(t, 192, generator=g).to(dev).half()
    ref = F.relu(F.linear(xin.float(), w.to(dev).half().float(), b.to(dev)) + add.float())
    for tag, r in (("tc", tc), ("ck", chk)):
        y = r.linear(L, xin, add0=add)
        torch.cuda.synchronize()
        e = _diff(y, ref)
        _report(test="linear", impl=tag, err=e)
        assert e[0] <= 4e-3, (tag, e)


@pytest.mark.parametrize("lens", [[192], [768, 192, 384], [100, 33, 64, 65]])
def test_attention_varlen(dev, runners, lens):
    tc, _ = runners
    d = 96
    g = torch.Generator().manual_seed(len(lens))
    t = sum(lens)
    qk = torch.randn(t, 2 * d, generator=g).to(dev).half()
    v = torch.randn(t, d, generator=g).to(dev).half()
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    scale = 1.0 / math.sqrt(d)
    out = tc.attention(qk[:, :d], qk[:, d:], v, cu, max(lens), scale)
    torch.cuda.synchronize()
    ref = torch.empty(t, d, device=dev)
    o = 0
    for n in lens:
        q, k, vv = qk[o:o + n, :d].float(), qk[o:o + n, d:].float(), v[o:o + n].float()
        ref[o:o + n] = torch.softmax(q @ k.t() * scale, dim=-1) @ vv
        o += n
    e = _diff(out, ref)
    _report(test="attention", lens=lens, err=e)
    assert e[0] <= 3e-3, e      # fp16 P and fp16 output rounding, |out| <~ 1


def test_layernorm_maxpool_stem_add(dev, runners):
    tc, _ = runners
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(1000, 96, generator=g) * 2 + 0.3).to(dev).half()
    pos = torch.randn(1000, 96, generator=g).to(dev).half()
    gamma = (torch.rand(96, generator=g) + 0.5).to(dev)
    beta = (torch.randn(96, generator=g) * 0.1).to(dev)
    y, y2 = tc.layernorm(x, gamma, beta, 1e-5, pos=pos)
    ref = F.layer_norm(x.float(), (96,), gamma, beta, 1e-5)
    assert _diff(y, ref)[0] <= 4e-3
    assert _diff(y2, ref + pos.float())[0] <= 8e-3
    # max-pool
    xm = torch.randn(3, 64, 48, 96, generator=g).to(dev).half()
    ym = tc.maxpool(xm)
    refm = F.max_pool2d(xm.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert tuple(ym.shape) == (3, 32, 24, 96) and _diff(ym, refm)[0] == 0.0
    # stem convs (Cin = 3 and 1), fp32 NCHW in -> fp16 NHWC out
    for cin in (3, 1):
        xs = torch.randn(2, cin, 64, 48, generator=g).to(dev)
        w = ((torch.rand(64, cin, 3, 3, generator=g) * 2 - 1) / math.sqrt(cin * 9)).to(dev)
        sc = (torch.rand(64, generator=g) + 0.5).to(dev)
        bi = (torch.randn(64, generator=g) * 0.1).to(dev)
        wk = w.permute(1, 2, 3, 0).reshape(-1, 64).contiguous()
        ys = tc.stem(xs, wk, sc, bi, 64)
        refs = F.relu(F.conv2d(xs, w, None, 2, 1) * sc.view(1, -1, 1, 1) + bi.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
        assert tuple(ys.shape) == (2, 32, 24, 64)
        assert _diff(ys, refs)[0] <= 4e-3
    a = torch.randn(4096, generator=g).to(dev).half()
    b = torch.randn(4096, generator=g).to(dev).half()
    assert _diff(tc.add(a, b), a.float() + b.float())[0] <= 4e-3
    torch.cuda.synchronize()


def test_bad_arguments_raise(dev, runners):
    from i2r_b200 import capi
    from i2r_b200.ops import ConvLayer
    tc, _ = runners
    with pytest.raises(ValueError):
        ConvLayer([torch.zeros(16, 40)], [0], [0], torch.ones(16), torch.zeros(16), device=dev)   # Cin % 48/64
    with pytest.raises(capi.I2RError):
        tc.attention(torch.zeros(8, 72, device=dev).half(), torch.zeros(8, 72, device=dev).half(),
                     torch.zeros(8, 72, device=dev).half(), torch.tensor([0, 8], dtype=torch.int32, device=dev),
                     8, 1.0)   # head dim 72 unsupported


def test_tma_residual_group(dev, runners):
    """BasicBlock second conv for three resolution branches in one persistent grid (residual addends,
    N-split halves writing channel slices)."""
    tc, chk = runners
    g = torch.Generator().manual_seed(17)
    specs_in = []
    for (c, h, w) in ((48, 64, 48), (96, 32, 24), (192, 16, 12)):
        L, wt, sc, bi = _mk_conv(c, c, 3, 1, True, dev, c)
        x = torch.randn(4, h, w, c, generator=g).to(dev).half()
        res = torch.randn(4, h, w, c, generator=g).to(dev).half()
        specs_in.append((L, wt, sc, bi, x, res))
    outs = tc.conv_group([(L, x, {"add0": res}) for L, _, _, _, x, res in specs_in])
    torch.cuda.synchronize()
    for (L, wt, sc, bi, x, res), y in zip(specs_in, outs):
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.to(dev).half().float(), None, 1, 1)
        ref = ref * sc.to(dev).view(1, -1, 1, 1) + bi.to(dev).view(1, -1, 1, 1)
        ref = F.relu(ref + res.float().permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        e = _diff(y, ref)
        _report(test="tma_group", c=L.cout, err=e)
        assert e[0] <= 8e-3, (L.cout, e)


# ---------------------------------------------------------------------------------------------- split-operand mode
SPLIT_CASES = [
    # name, NB, H, W, Cin, Cout, k, stride, relu, residual
    ("s_c48_3x3", 2, 64, 48, 48, 48, 3, 1, True, True),           # halo kernel, resident weights
    ("s_c192_3x3", 3, 16, 12, 192, 192, 3, 1, True, True),        # halo kernel, streamed weights (9 K-chunks)
    ("s_c96_3x3_s2", 2, 32, 24, 96, 192, 3, 2, True, False),      # gather kernel
    ("s_c64to256_1x1", 1, 64, 48, 64, 256, 1, 1, False, False),   # halo 1x1
    ("s_linear_96to192", 1, 3072, 1, 96, 192, 1, 1, True, False),
]


@pytest.mark.parametrize("case", SPLIT_CASES, ids=[c[0] for c in SPLIT_CASES])
def test_split_operand_conv(case, dev, runners, gather_runner):
    """Pair tensors (hi | lo) in, pair tensor out: the result must match an fp32 torch convolution on the FULL
    precision values to ~1e-5 (three-term products), two orders of magnitude below the fp16 path."""
    from i2r_b200.ops import ConvLayer
    from i2r_b200.packing import conv_taps, merge_pair, split_pair
    name, nb, h, w, cin, cout, k, stride, relu, residual = case
    tc, chk = runners
    g = torch.Generator().manual_seed(hash(name) % 1000)
    wt = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) / math.sqrt(cin * k * k)
    scale = torch.rand(cout, generator=g) + 0.5
    bias = torch.randn(cout, generator=g) * 0.1
    mats, dys, dxs = conv_taps(wt, pad=k // 2)
    L = ConvLayer(mats, dys, dxs, scale, bias, stride=stride, relu=relu, device=dev, split=True)
    x32 = torch.randn(nb, h, w, cin, generator=g)
    x = split_pair(x32).to(dev)
    oh, ow = (h + stride - 1) // stride, (w + stride - 1) // stride
    kw = {}
    ref = F.conv2d(merge_pair(x.cpu()).permute(0, 3, 1, 2).double(), wt.double(), None, stride, k // 2)
    ref = ref * scale.double().view(1, -1, 1, 1) + bias.double().view(1, -1, 1, 1)
    if residual:
        r32 = torch.randn(nb, oh, ow, cout, generator=g)
        res = split_pair(r32).to(dev)
        kw["add0"] = res
        ref = ref + merge_pair(res.cpu()).permute(0, 3, 1, 2).double()
    if relu:
        ref = F.relu(ref)
    ref = ref.permute(0, 2, 3, 1).float()
    errs = {}
    for tag, r in (("tc", tc), ("ck", chk), ("ga", gather_runner)):
        y = r.conv(L, x, **kw)
        torch.cuda.synchronize()
        assert tuple(y.shape) == (nb, oh, ow, 2 * cout)
        errs[tag] = float((merge_pair(y.cpu()) - ref).abs().max())
    _report(test="split_conv", case=name, err=errs)
    assert all(v <= 6e-5 for v in errs.values()), errs   # fp32 accumulation over K up to 1728 + dropped lo*lo terms


def test_split_small_kernels(dev, runners):
    from i2r_b200.packing import merge_pair, split_pair
    tc, _ = runners
    tc.split = True
    try:
        g = torch.Generator().manual_seed(11)
        # layernorm (+pos), add, maxpool on pair tensors
        x32 = torch.randn(1000, 96, generator=g) * 3
        p32 = torch.randn(1000, 96, generator=g)
        gamma, beta = torch.rand(96, generator=g) + 0.5, torch.randn(96, generator=g) * 0.1
        y, y2 = tc.layernorm(split_pair(x32).to(dev), gamma.to(dev), beta.to(dev), pos=split_pair(p32).to(dev))
        ref = F.layer_norm(merge_pair(split_pair(x32)), (96,), gamma, beta, 1e-5)
        e_ln = float((merge_pair(y.cpu()) - ref).abs().max())
        e_ln2 = float((merge_pair(y2.cpu()) - (ref + merge_pair(split_pair(p32)))).abs().max())
        a32, b32 = torch.randn(4, 8, 6, 48, generator=g), torch.randn(4, 8, 6, 48, generator=g)
        s = tc.add(split_pair(a32).to(dev), split_pair(b32).to(dev))
        e_add = float((merge_pair(s.cpu()) - (merge_pair(split_pair(a32)) + merge_pair(split_pair(b32)))).abs().max())
        m32 = torch.randn(3, 17, 11, 96, generator=g)
        mp = tc.maxpool(split_pair(m32).to(dev))
        refm = F.max_pool2d(merge_pair(split_pair(m32)).permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
        e_mp = float((merge_pair(mp.cpu()) - refm).abs().max())
        # stem
        xs = torch.randn(2, 3, 64, 48, generator=g)
        w = torch.randn(64, 3, 3, 3, generator=g) * 0.2
        sc, bi = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
        ys = tc.stem(xs.to(dev), w.permute(1, 2, 3, 0).reshape(-1, 64).contiguous().to(dev), sc.to(dev), bi.to(dev), 64)
        refs = F.relu(F.conv2d(xs, w, None, 2, 1) * sc.view(1, -1, 1, 1) + bi.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
        e_st = float((merge_pair(ys.cpu()) - refs).abs().max())
        torch.cuda.synchronize()
    finally:
        tc.split = False
    _report(test="split_small", ln=e_ln, ln_pos=e_ln2, add=e_add, maxpool=e_mp, stem=e_st)
    assert e_ln <= 1e-5 and e_ln2 <= 1e-5 and e_add <= 2e-6 and e_mp <= 1e-6 and e_st <= 2e-5, (e_ln, e_ln2, e_add, e_mp, e_st)


@pytest.mark.parametrize("lens", [[3072], [192, 768, 64]])
def test_split_attention(dev, runners, lens):
    from i2r_b200.packing import merge_pair, split_pair
    tc, _ = runners
    g = torch.Generator().manual_seed(13)
    t, d = sum(lens), 96
    qk32 = torch.randn(t, 2 * d, generator=g)
    v32 = torch.randn(t, d, generator=g)
    qk = split_pair(qk32).to(dev)          # [t, (q_hi k_hi | q_lo k_lo)]
    v = split_pair(v32).to(dev)            # [t, (v_hi | v_lo)]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32).to(dev)
    scale = d ** -0.5
    out = tc.attention(qk[:, :d], qk[:, d:2 * d], v[:, :d], cu, max(lens), scale, lo=(2 * d, 2 * d, d))
    torch.cuda.synchronize()
    qf = merge_pair(qk.cpu())
    vf = merge_pair(v.cpu())
    ref = torch.empty(t, d)
    o = 0
    for n in lens:
        s = torch.softmax(qf[o:o + n, :d].double() @ qf[o:o + n, d:].double().t() * scale, dim=-1)
        ref[o:o + n] = (s @ vf[o:o + n].double()).float()
        o += n
    err = float((merge_pair(out.cpu()) - ref).abs().max())
    _report(test="split_attention", lens=lens, err=err)
    # fp16 probabilities bound the error: ~2^-11 relative on each of ~n averaged terms
    assert tuple(out.shape) == (t, 2 * d) and err <= 4e-4, err
