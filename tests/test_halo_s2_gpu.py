"""Stride-2 3x3 convolutions on `conv_halo_kernel` (parity-plane staging through TMA element strides; csrc/conv_halo.cu
header, profiles/r02_tma_element_stride_probe.txt) against float64 torch on the operands the kernel sees.  Shapes are the
reference's: the second stem convolution (lib/models/interformer_pureMulti.py:680-682), transition1 / transition2
(:543-582), the stride-2 chains of the fuse layers incl. identity and up-sampled addends (:353-387, :392-410)."""
import pytest
import torch
import torch.nn.functional as F

import paths  # noqa: F401
from i2r_b200 import capi
from i2r_b200.ops import ConvLayer, Runner, split_precision
from i2r_b200.packing import conv_taps, merge_pair, split_pair

pytestmark = pytest.mark.gpu

# (cin, cout, crops, IH, IW, add0, add1_shift or None, relu)
CASES = {
    "stem_conv2_64_64": (64, 64, 4, 128, 96, False, None, True),
    "transition1_256_96": (256, 96, 4, 64, 48, False, None, True),
    "fuse_hop_48_48": (48, 48, 6, 64, 48, False, None, True),
    "fuse_96_192_two_addends": (96, 192, 5, 32, 24, True, 0, True),
    "fuse_48_96_upsampled_addend": (48, 96, 5, 64, 48, True, 1, True),
    "odd_size_48_48": (48, 48, 3, 50, 38, True, None, False),
    "tiny_192_192": (192, 192, 2, 16, 12, False, None, True),
}


def _pair(t, split, dev):
    return (split_pair(t) if split else t.half()).to(dev)


def _val(t, split):
    return (merge_pair(t.cpu()) if split else t.cpu().float()).double()


@pytest.mark.parametrize("split", [False, True], ids=["fp16", "split"])
@pytest.mark.parametrize("case", list(CASES))
def test_stride2_conv_on_halo_kernel(case, split):
    cin, cout, nb, ih, iw, has0, sh1, relu = CASES[case]
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(len(case) * 7 + cin)
    w = (torch.rand(cout, cin, 3, 3, generator=g) * 2 - 1) * (3.0 / (9 * cin)) ** 0.5
    scale = torch.rand(cout, generator=g) * 0.5 + 0.75
    bias = torch.rand(cout, generator=g) - 0.5
    mats, dys, dxs = conv_taps(w, pad=1)
    with split_precision(split):
        L = ConvLayer(mats, dys, dxs, scale, bias, stride=2, relu=relu, device=dev)
    oh, ow = (ih + 1) // 2, (iw + 1) // 2
    x32 = torch.randn(nb, ih, iw, cin, generator=g)
    a0_32 = torch.randn(nb, oh, ow, cout, generator=g) if has0 else None
    a1_32 = torch.randn(nb, oh >> sh1, ow >> sh1, cout, generator=g) if sh1 is not None else None
    x = _pair(x32, split, dev)
    kw = {}
    if has0:
        kw["add0"] = _pair(a0_32, split, dev)
    if sh1 is not None:
        kw["add1"], kw["add1_shift"] = _pair(a1_32, split, dev), sh1
    r = Runner(dev, 0)
    r.split = split
    lib = capi.load()
    probs, _ = r.problems(L, x, **dict(kw))
    if split and cout > 96:
        # split-operand layers are not N-split: the 3-tap weight slots of > 96 channels do not fit beside two 71 KB stages
        assert not any(lib.i2r_conv_halo_supported(p) for p in probs)
    else:
        assert all(lib.i2r_conv_halo_supported(p) for p in probs), "must run on conv_halo_kernel"
    out = r.conv(L, x, **dict(kw))
    torch.cuda.synchronize()
    assert tuple(out.shape) == (nb, oh, ow, (2 if split else 1) * cout)
    for _ in range(20):
        assert torch.equal(r.conv(L, x, **dict(kw)), out)
    torch.cuda.synchronize()

    def q(t):       # weights as the kernel holds them: BN scale folded in fp32, then fp16 (hi + lo in split mode)
        hi = t.half().double()
        return hi + (t.double() - hi).half().double() if split else hi
    wq = q(w * scale.view(-1, 1, 1, 1))
    ref = F.conv2d(_val(x, split).permute(0, 3, 1, 2), wq, stride=2, padding=1) + bias.double().view(1, -1, 1, 1)
    if has0:
        ref = ref + _val(kw["add0"], split).permute(0, 3, 1, 2)
    if sh1 is not None:
        a1 = _val(kw["add1"], split).permute(0, 3, 1, 2)
        if sh1:
            a1 = a1.repeat_interleave(1 << sh1, 2).repeat_interleave(1 << sh1, 3)
        ref = ref + a1
    if relu:
        ref = ref.clamp_min(0)
    got = _val(out, split).permute(0, 3, 1, 2)
    err = float((got - ref).abs().max())
    assert err <= (3e-5 if split else 6e-3) * max(1.0, float(ref.abs().max())), err


def test_vanilla_forward_has_no_stride2_launch_on_the_gather_kernel():
    """After this change only the transposed convolutions (4 phase problems of 2x2 taps) are left on igemm_tc_kernel in the
    vanilla model: every launch of an eager forward is classified by the kernel that takes it."""
    from helpers import build_model, inputs_for
    _, model, _ = build_model()
    model = model.cuda()
    model.use_cuda_graph = False
    model.check_impl = False
    model.prepare("cuda:0")
    r = model._program.runner
    lib = capi.load()
    seen = {"halo_s2": 0, "igemm_s2": 0, "igemm_other": 0}
    orig = r._launch_now

    def spy(problems):
        import ctypes
        for p in problems:
            halo = bool(lib.i2r_conv_halo_supported(ctypes.byref(p)))
            if p.stride == 2:
                seen["halo_s2" if halo else "igemm_s2"] += 1
            elif not halo:
                seen["igemm_other"] += 1
        return orig(problems)
    r._launch_now = spy
    length = [2, 1]
    x, pm = inputs_for(length)
    with torch.no_grad():
        model(x, pm, length)
    torch.cuda.synchronize()
    assert seen["igemm_s2"] == 0 and seen["halo_s2"] >= 12, seen
    assert 0 < seen["igemm_other"] <= 8, seen      # two ConvTranspose2d applications x four phase problems
