"""Routing rules of the halo kernel (`i2r_conv_halo_supported`, include/i2r.h) -- pure host logic of the C library, no
GPU needed: which problems of the reference's layer zoo run on `conv_halo_kernel` and which fall back to the gather kernel.
Problems are built by the product's own `Runner.problem` / `Runner.problems` on CPU tensors."""
import ctypes

import pytest
import torch

import paths  # noqa: F401
from i2r_b200 import capi
from i2r_b200.ops import ConvLayer, Runner, split_precision
from i2r_b200.packing import conv_taps


@pytest.fixture(scope="module")
def lib():
    return capi.load()


def _runner():
    r = Runner.__new__(Runner)
    r.lib, r.device, r.impl, r.use_tma = capi.load(), torch.device("cpu"), 0, True
    r.launches, r.split, r.timing = 0, False, None
    r._chain, r.chain_enabled, r._in_parallel, r.chains = None, False, 0, 0
    return r


def _layer(cin, cout, k, stride=1, split=False):
    w = torch.randn(cout, cin, k, k) * 0.05
    mats, dys, dxs = conv_taps(w, pad=k // 2)
    with split_precision(split):
        return ConvLayer(mats, dys, dxs, torch.ones(cout), torch.zeros(cout), stride=stride, relu=True, device="cpu")


def _x(nb, h, w, c, split=False):
    return torch.zeros(nb, h, w, (2 if split else 1) * c, dtype=torch.float16)


def _ok(lib, p):
    return bool(lib.i2r_conv_halo_supported(ctypes.byref(p)))


def test_stride1_and_1x1_problems_are_halo_problems(lib):
    r = _runner()
    for cin, cout, k in ((48, 48, 3), (96, 96, 3), (64, 256, 1), (256, 64, 1), (96, 17, 1)):
        p, _ = r.problem(_layer(cin, cout, k), _x(2, 32, 24, cin))
        assert _ok(lib, p), (cin, cout, k)


def test_stride2_rules(lib):
    """Stride-2 3x3 (parity planes): supported while two 71 KB activation stages and the weights (resident, else a ring of two
    three-tap slots) fit; wider layers run as two halves; split-operand layers keep the three-pass K layout."""
    r = _runner()
    for cin, cout in ((64, 64), (256, 96), (48, 48), (48, 96), (64, 96)):
        p, out = r.problem(_layer(cin, cout, 3, 2), _x(2, 64, 48, cin))
        assert _ok(lib, p) and tuple(out.shape) == (2, 32, 24, cout), (cin, cout)
    p, _ = r.problem(_layer(96, 192, 3, 2), _x(2, 32, 24, 96))
    assert not _ok(lib, p)                                   # one 192-channel problem: slots too large ...
    probs, out = r.problems(_layer(96, 192, 3, 2), _x(2, 32, 24, 96))
    assert len(probs) == 2 and all(_ok(lib, q) for q in probs) and tuple(out.shape) == (2, 16, 12, 192)   # ... two halves fit
    p, _ = r.problem(_layer(256, 96, 3, 2, split=True), _x(2, 64, 48, 256, split=True))
    assert _ok(lib, p)
    p, _ = r.problem(_layer(96, 192, 3, 2, split=True), _x(2, 32, 24, 96, split=True))
    assert not _ok(lib, p)
    # odd input sizes: OH = ceil(IH / 2)
    p, out = r.problem(_layer(48, 48, 3, 2), _x(1, 25, 19, 48))
    assert _ok(lib, p) and tuple(out.shape) == (1, 13, 10, 48)


def test_addend_and_resampling_rules(lib):
    r = _runner()
    L = _layer(48, 96, 3, 2)
    x = _x(2, 64, 48, 48)
    a0, a1 = _x(2, 32, 24, 96), _x(2, 16, 12, 96)
    p, _ = r.problem(L, x, add0=a0, add1=a1, add1_shift=1)
    assert _ok(lib, p)                                       # HRNet fuse: identity + up-sampled lower-resolution term
    p, _ = r.problem(L, x, add0=a1, add0_shift=1)
    assert not _ok(lib, p)                                   # only the SECOND addend may be up-sampled
    p, _ = r.problem(_layer(48, 96, 3, 2), _x(2, 50, 38, 48), add0=_x(2, 25, 19, 96), add1=_x(2, 12, 9, 96), add1_shift=1)
    assert not _ok(lib, p)                                   # odd output size: no exact 2x up-sampling
    # nearest-upsampled INPUT (in_shift) and transposed-convolution phases (2x2 taps, out_mul 2) stay on the gather kernel
    p, _ = r.problem(_layer(96, 48, 1), _x(2, 16, 12, 96), in_shift=1)
    assert not _ok(lib, p)
    w = torch.randn(96, 96, 2, 2) * 0.05
    mats = [w[:, :, i, j] for i in range(2) for j in range(2)]
    Ld = ConvLayer(mats, [0, 0, 1, 1], [0, 1, 0, 1], torch.ones(96), torch.zeros(96), device="cpu")
    p, _ = r.problem(Ld, _x(2, 16, 12, 96), out_mul=2, out_off=(0, 0))
    assert not _ok(lib, p)


def test_chain_entry_point_argument_checks_need_no_device(lib):
    r = _runner()
    L = _layer(48, 48, 3)
    p, _ = r.problem(L, _x(5, 64, 48, 48))
    arr = (capi.ConvProblem * 2)(p, p)
    assert lib.i2r_conv_halo_chain_workspace(arr, 2) == 2 * 5 * 4
    counts = (ctypes.c_int * 1)(2)
    assert lib.i2r_conv_halo_chain(arr, counts, 1, None, 0, None) == -1          # no workspace
    assert lib.i2r_conv_halo_chain(arr, counts, capi.I2R_MAX_CHAIN_LAYERS + 1, None, 0, None) == -1
