"""Host logic of chained halo launches (`Runner.chain`, i2r_b200/ops.py) against a recording stand-in for the C library:
which launches are deferred into a chain, what flushes it, stream-order of mixed sequences, the fall-back when the
entry point refuses.  (The device side is covered by tests/test_halo_chain_gpu.py.)"""
import pytest
import torch

import paths  # noqa: F401
from i2r_b200 import capi, ops
from i2r_b200.ops import ConvLayer, Runner
from i2r_b200.packing import conv_taps


class RecordingLib:
    def __init__(self, chain_rc=0):
        self.calls = []
        self.chain_rc = chain_rc

    def i2r_conv_halo_supported(self, p):
        return 1 if p._obj.stride == 1 else 0

    def i2r_conv_halo_chain_workspace(self, arr, n):
        return 4 * sum(arr[i].NB for i in range(n))

    def i2r_conv_halo_chain(self, arr, counts, nlayers, ws, ws_bytes, stream):
        self.calls.append(("chain", [counts[i] for i in range(nlayers)]))
        return self.chain_rc

    def i2r_conv_halo(self, arr, n, stream):
        self.calls.append(("halo", n))
        return 0

    def i2r_conv_igemm(self, arr, n, impl, stream):
        self.calls.append(("igemm", n))
        return 0

    def i2r_add_f16(self, *a):
        self.calls.append(("add",))
        return 0

    def i2r_last_error(self):
        return b"stub"


def _runner(lib, monkeypatch):
    r = Runner.__new__(Runner)
    r.lib, r.device, r.impl, r.use_tma = lib, torch.device("cpu"), 0, True
    r.launches, r.split, r.timing = 0, False, None
    r._side_streams, r.concurrent_branches = [], False
    r.chain_enabled, r._chain, r._chain_ws, r._in_parallel, r.chains = True, None, None, 0, 0
    monkeypatch.setattr(ops, "_stream_ptr", lambda: None)
    return r


def _layer(cin, cout, k, stride=1):
    w = torch.randn(cout, cin, k, k) * 0.1
    mats, dys, dxs = conv_taps(w, pad=k // 2)
    return ConvLayer(mats, dys, dxs, torch.ones(cout), torch.zeros(cout), stride=stride, relu=True, device="cpu")


def test_consecutive_halo_launches_become_one_chain(monkeypatch):
    lib = RecordingLib()
    r = _runner(lib, monkeypatch)
    L = _layer(48, 48, 3)
    x = torch.zeros(2, 16, 8, 48, dtype=torch.float16)
    with r.chain():
        h = r.conv(L, x)
        h = r.conv_group([(L, h, {}), (L, x, {})])[0]
        h = r.conv(L, h, add0=x)
        assert lib.calls == []                       # everything deferred
    assert lib.calls == [("chain", [1, 2, 1])]
    assert r.launches == 1 and r.chains == 1


def test_other_ops_and_unsupported_problems_flush_in_stream_order(monkeypatch):
    lib = RecordingLib()
    r = _runner(lib, monkeypatch)
    L, Ls2 = _layer(48, 48, 3), _layer(48, 48, 3, stride=2)
    x = torch.zeros(2, 16, 8, 48, dtype=torch.float16)
    with r.chain():
        a = r.conv(L, x)
        b = r.conv(L, a)
        c = r.add(a, b)                              # not a conv: the two deferred layers go first
        d = r.conv(L, c)                             # a single deferred layer is a plain launch
        e = r.conv(Ls2, d)                           # stride 2 is not a halo problem: flush, then the gather kernel
        f = r.conv(L, e)
        g = r.conv(L, f)
    assert lib.calls == [("chain", [1, 1]), ("add",), ("halo", 1), ("igemm", 1), ("chain", [1, 1])]


def test_refused_chain_falls_back_to_single_launches(monkeypatch):
    lib = RecordingLib(chain_rc=capi.E_UNSUPPORTED)
    r = _runner(lib, monkeypatch)
    L = _layer(48, 48, 3)
    x = torch.zeros(2, 16, 8, 48, dtype=torch.float16)
    with r.chain():
        h = r.conv(L, x)
        h = r.conv(L, h)
        h = r.conv(L, h)
    assert lib.calls == [("chain", [1, 1, 1]), ("halo", 1), ("halo", 1), ("halo", 1)]
    assert r.chains == 0 and r.launches == 3


def test_chain_limits_split_long_sequences_and_disabled_runner_is_passthrough(monkeypatch):
    lib = RecordingLib()
    r = _runner(lib, monkeypatch)
    L = _layer(48, 48, 3)
    x = torch.zeros(1, 16, 8, 48, dtype=torch.float16)
    with r.chain():
        h = x
        for _ in range(capi.I2R_MAX_CHAIN_LAYERS + 3):
            h = r.conv(L, h)
    assert lib.calls == [("chain", [1] * capi.I2R_MAX_CHAIN_LAYERS), ("chain", [1] * 3)]
    lib.calls.clear()
    r.chain_enabled = False
    with r.chain():
        h = r.conv(L, x)
        h = r.conv(L, h)
    assert lib.calls == [("halo", 1), ("halo", 1)]
    # inside concurrent branches (Runner.parallel) a chain context is inert as well
    r.chain_enabled = True
    r._in_parallel = 1
    lib.calls.clear()
    with r.chain():
        r.conv(L, x)
        r.conv(L, x)
    assert lib.calls == [("halo", 1), ("halo", 1)]


def test_error_inside_chain_context_drops_the_deferred_layers(monkeypatch):
    lib = RecordingLib()
    r = _runner(lib, monkeypatch)
    L = _layer(48, 48, 3)
    x = torch.zeros(1, 16, 8, 48, dtype=torch.float16)
    with pytest.raises(ValueError):
        with r.chain():
            r.conv(L, x)
            raise ValueError("boom")
    assert r._chain is None and lib.calls == []
