"""GPU parity of the tcgen05 attention kernel (i2r_attention_tc) and of the transposed (V^T) GEMM output that feeds
it, against torch float64 references of the same operation (nn.MultiheadAttention(nhead=1) arithmetic:
torch/nn/functional.py:6632-6650 as called from lib/models/transpose_h.py:165-240, lib/models/attention.py:68-73)."""
import json
import math
import os

import pytest
import torch

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


def _ref(q, k, v, lens, scale):
    out = torch.empty(q.shape[0], v.shape[1], dtype=torch.float64)
    o = 0
    for n in lens:
        s = torch.softmax(q[o:o + n].double() @ k[o:o + n].double().t() * scale, dim=-1)
        out[o:o + n] = s @ v[o:o + n].double()
        o += n
    return out


# 3072 = one crop of the intra-human stage; 768 / 192 = inter-human sequences (4 / 1 persons); 200, 72 and 1000 exercise
# partially filled key blocks, a single-tile CTA and sequences that do not start on a 128-token boundary
LENS = [[3072], [768] * 4, [192, 768, 384], [200, 72, 1000, 128, 256], [3072, 3072, 192]]


@pytest.mark.parametrize("lens", LENS)
def test_attention_tc_fp16(dev, tc, lens):
    d = 96
    g = torch.Generator().manual_seed(sum(lens))
    t = sum(lens)
    qk = torch.randn(t, 2 * d, generator=g).half()
    v = torch.randn(t, d, generator=g).half()
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    scale = 1.0 / math.sqrt(d)
    qkd = qk.to(dev)
    vt = v.t().contiguous().to(dev)
    out = tc.attention_tc(qkd[:, :d], qkd[:, d:], vt, cu, max(lens), scale)
    torch.cuda.synchronize()
    ref = _ref(qk[:, :d], qk[:, d:], v, lens, scale)
    err = float((out.cpu().double() - ref).abs().max())
    _report(test="attention_tc_fp16", lens=lens, err=err)
    assert tuple(out.shape) == (t, d) and err <= 3e-3, err      # fp16 P and fp16 output rounding, |out| <~ 1


@pytest.mark.parametrize("lens", LENS)
def test_attention_tc_split(dev, tc, lens):
    from i2r_b200.packing import merge_pair, split_pair
    d = 96
    g = torch.Generator().manual_seed(7 + sum(lens))
    t = sum(lens)
    q32, k32, v32 = (torch.randn(t, d, generator=g) for _ in range(3))
    q, k = split_pair(q32), split_pair(k32)             # [t, (hi | lo)]
    vt = split_pair(v32)                                # [t, (hi | lo)] -> rows hi channels then lo channels
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    scale = 1.0 / math.sqrt(d)
    out = tc.attention_tc(q.to(dev), k.to(dev), vt.t().contiguous().to(dev), cu, max(lens), scale, split=True)
    torch.cuda.synchronize()
    ref = _ref(merge_pair(q), merge_pair(k), merge_pair(vt), lens, scale)
    err = float((merge_pair(out.cpu()).double() - ref).abs().max())
    _report(test="attention_tc_split", lens=lens, err=err)
    # fp16 probabilities bound the error: ~2^-11 relative on each of ~n averaged terms
    assert tuple(out.shape) == (t, 2 * d) and err <= 4e-4, err


def test_attention_tc_large_logits(dev, tc):
    """Row maxima that keep growing along the keys force the lazy O / l rescaling path."""
    d, lens = 96, [1024, 640]
    g = torch.Generator().manual_seed(5)
    t = sum(lens)
    q = torch.randn(t, d, generator=g).half()
    k = torch.randn(t, d, generator=g)
    k = (k * torch.linspace(0.5, 6.0, t).unsqueeze(1)).half()      # later keys score higher
    v = torch.randn(t, d, generator=g).half()
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    scale = 1.0 / math.sqrt(d)
    out = tc.attention_tc(q.to(dev), k.to(dev), v.t().contiguous().to(dev), cu, max(lens), scale)
    torch.cuda.synchronize()
    ref = _ref(q, k, v, lens, scale)
    err = float((out.cpu().double() - ref).abs().max())
    _report(test="attention_tc_large_logits", err=err)
    assert err <= 6e-3, err


@pytest.mark.parametrize("split", [False, True])
def test_linear_transposed_output(dev, tc, split):
    """I2R_F_OUT_T16: the V projection written channel-major (V^T), the B operand of the P.V product."""
    from i2r_b200.ops import ConvLayer, split_precision
    from i2r_b200.packing import merge_pair, split_pair
    g = torch.Generator().manual_seed(3)
    t, cin, cout = 1000, 96, 96
    w = (torch.rand(cout, cin, generator=g) * 2 - 1) / math.sqrt(cin)
    b = torch.randn(cout, generator=g) * 0.1
    x32 = torch.randn(t, cin, generator=g)
    with split_precision(split):
        L = ConvLayer([w], [0], [0], torch.ones(cout), b, device=dev)
    x = (split_pair(x32) if split else x32.half()).to(dev)
    p, out = tc.linear_problem(L, x, out_mode="t16")
    tc.launch([p])
    torch.cuda.synchronize()
    xin = merge_pair(x.cpu()).double() if split else x.cpu().double()
    ref = xin @ w.double().t() + b.double()
    got = out.cpu()
    assert tuple(got.shape) == ((2 if split else 1) * cout, t)
    val = merge_pair(got.t().contiguous()).double() if split else got.t().double()
    err = float((val - ref).abs().max())
    _report(test="linear_t16", split=split, err=err)
    assert err <= (2e-5 if split else 4e-3), err


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("t,with_pos", [(1000, True), (128, False), (6144, True)])
def test_encoder_tail(dev, tc, split, t, with_pos):
    """Fused out-proj + residual + LN1 + FFN + residual + LN2 (+ pos) vs float64 torch on the same operands
    (TransformerEncoderLayer.forward_post after self_attn, lib/models/transpose_h.py:205-222)."""
    import torch.nn.functional as F
    from i2r_b200.ops import EncoderTailParams
    from i2r_b200.packing import merge_pair, split_pair
    g = torch.Generator().manual_seed(11 + t)
    d, f = 96, 192

    def u(*shape, s=1.0):
        return (torch.rand(*shape, generator=g) * 2 - 1) * s
    w_out, b_out = u(d, d, s=d ** -0.5), u(d, s=0.1)
    w1, b1 = u(f, d, s=d ** -0.5), u(f, s=0.1)
    w2, b2 = u(d, f, s=f ** -0.5), u(d, s=0.1)
    g1, be1, g2, be2 = 1 + u(d, s=0.3), u(d, s=0.1), 1 + u(d, s=0.3), u(d, s=0.1)
    attn32, src32, pos32 = (torch.randn(t, d, generator=g) for _ in range(3))
    enc = split_pair if split else (lambda v: v.half())
    dec = (lambda v: merge_pair(v).double()) if split else (lambda v: v.double())
    attn, src, pos = enc(attn32), enc(src32), enc(pos32)
    tail = EncoderTailParams(w_out, b_out, w1, b1, w2, b2, g1, be1, g2, be2, device=dev, split=split)
    out, out_pos = tc.encoder_tail(tail, attn.to(dev), src.to(dev), pos=pos.to(dev) if with_pos else None)
    torch.cuda.synchronize()
    D = torch.float64
    s1 = F.layer_norm(dec(attn) @ w_out.to(D).t() + b_out.to(D) + dec(src), (d,), g1.to(D), be1.to(D), 1e-5)
    h = F.relu(s1 @ w1.to(D).t() + b1.to(D))
    ref = F.layer_norm(h @ w2.to(D).t() + b2.to(D) + s1, (d,), g2.to(D), be2.to(D), 1e-5)
    err = float((dec(out.cpu()) - ref).abs().max())
    err_pos = float((dec(out_pos.cpu()) - (ref + dec(pos))).abs().max()) if with_pos else 0.0
    _report(test="encoder_tail", split=split, t=t, err=err, err_pos=err_pos)
    tol = 3e-5 if split else 1.5e-2      # fp16 weights / activations through three GEMMs and two LayerNorms
    assert (out_pos is None) == (not with_pos)
    assert err <= tol and err_pos <= 2 * tol, (err, err_pos)


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("cin,h,w", [(3, 256, 192), (1, 256, 192), (3, 384, 288), (1, 72, 40)])
def test_stem_tc(dev, tc, split, cin, h, w):
    """Tensor-core stem (conv 3x3 s2 + folded BN + ReLU, fp32 NCHW -> fp16 NHWC) vs float64 torch; also partial tiles."""
    import torch.nn.functional as F
    from i2r_b200.packing import merge_pair
    g = torch.Generator().manual_seed(100 + cin + h)
    nb = 3
    x = torch.randn(nb, cin, h, w, generator=g) * 1.5
    wt = (torch.rand(64, cin, 3, 3, generator=g) * 2 - 1) / math.sqrt(cin * 9)
    sc = torch.rand(64, generator=g) + 0.5
    bi = torch.randn(64, generator=g) * 0.1
    wk = wt.permute(1, 2, 3, 0).reshape(-1, 64).contiguous()
    tc.split = split
    try:
        y = tc.stem(x.to(dev), wk.to(dev), sc.to(dev), bi.to(dev), 64)
        torch.cuda.synchronize()
    finally:
        tc.split = False
    ref = F.relu(F.conv2d(x.double(), wt.double(), None, 2, 1) * sc.double().view(1, -1, 1, 1)
                 + bi.double().view(1, -1, 1, 1)).permute(0, 2, 3, 1)
    got = merge_pair(y.cpu()).double() if split else y.cpu().double()
    err = float((got - ref).abs().max())
    _report(test="stem_tc", split=split, cin=cin, h=h, w=w, err=err, ref_max=float(ref.abs().max()))
    assert tuple(y.shape) == (nb, h // 2, w // 2, 128 if split else 64)
    assert err <= (2e-5 if split else 4e-3), err       # fp16 output rounding dominates the single-precision mode


@pytest.mark.parametrize("split", [False, True])
def test_attention_tc_head_dim_80(dev, tc, split):
    """Head dim 80 = d_model 78 of the HRFormer-B inter-human stage padded to the K=16 step (pad channels are zero)."""
    from i2r_b200.packing import merge_pair, split_pair
    d, real, lens = 80, 78, [1536, 296, 192]      # token totals must be multiples of 8 (V^T row stride)
    g = torch.Generator().manual_seed(80)
    t = sum(lens)
    q32, k32, v32 = (torch.randn(t, d, generator=g) for _ in range(3))
    for m in (q32, k32, v32):
        m[:, real:] = 0
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    scale = 1.0 / math.sqrt(real)
    if split:
        q, k, v = split_pair(q32), split_pair(k32), split_pair(v32)
        out = tc.attention_tc(q.to(dev), k.to(dev), v.t().contiguous().to(dev), cu, max(lens), scale, split=True)
        torch.cuda.synchronize()
        got, ref = merge_pair(out.cpu()).double(), _ref(merge_pair(q), merge_pair(k), merge_pair(v), lens, scale)
    else:
        q, k, v = q32.half(), k32.half(), v32.half()
        out = tc.attention_tc(q.to(dev), k.to(dev), v.t().contiguous().to(dev), cu, max(lens), scale)
        torch.cuda.synchronize()
        got, ref = out.cpu().double(), _ref(q, k, v, lens, scale)
    err = float((got - ref).abs().max())
    _report(test="attention_tc_d80", split=split, err=err)
    assert err <= (4e-4 if split else 3e-3), err
    assert float(got[:, real:].abs().max()) == 0.0
