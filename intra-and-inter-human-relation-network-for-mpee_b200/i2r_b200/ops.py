"""Op wrappers: fill i2r_conv_problem structs / argument lists and launch through the C ABI on
torch's current CUDA stream.  Activations are fp16 NHWC torch tensors (device memory plumbing only)."""
import ctypes

import torch

from . import capi
import contextlib

from .packing import ceil_to, pad_vec, pick_kc, pack_taps, pack_folded, split_virtual_taps


_SPLIT_DEFAULT = [False]
_CHANNEL_PAD = [0]


@contextlib.contextmanager
def channel_padding(multiple=16):
    """Layers constructed inside this context zero-pad their input channels (and output channels, except small heads)
    to a multiple of `multiple`: the K = 16 granularity of the tensor-core GEMMs for the 78/156/312/624-channel
    HRFormer-B branches.  Pad output channels get zero weights, scale 1 and bias 0, so they stay exactly zero."""
    old = _CHANNEL_PAD[0]
    _CHANNEL_PAD[0] = int(multiple)
    try:
        yield
    finally:
        _CHANNEL_PAD[0] = old


def pad_channels(c, small_ok=True):
    m = _CHANNEL_PAD[0]
    if not m or (small_ok and c <= 32):      # heatmap heads (17 / 14 joints) keep their channel count
        return c
    return (c + m - 1) // m * m


@contextlib.contextmanager
def split_precision(enabled=True):
    """Layers constructed inside this context use the split-operand mode (see packing.split_virtual_taps)."""
    old = _SPLIT_DEFAULT[0]
    _SPLIT_DEFAULT[0] = bool(enabled)
    try:
        yield
    finally:
        _SPLIT_DEFAULT[0] = old


class ConvLayer:
    """Device-resident parameters of one implicit-GEMM problem (weights packed, BN folded)."""

    def __init__(self, mats, dys, dxs, scale, bias, stride=1, relu=False, device="cuda", split=None):
        cout, cin = mats[0].shape
        cop, cip = pad_channels(cout), pad_channels(cin, small_ok=False)
        if (cop, cip) != (cout, cin):
            padded = []
            for m in mats:
                t = torch.zeros(cop, cip, dtype=torch.float32)
                t[:cout, :cin] = m.float()
                padded.append(t)
            mats = padded
            scale = torch.cat([scale.float().reshape(-1)[:cout], torch.ones(cop - cout)])
            bias = torch.cat([bias.float().reshape(-1)[:cout], torch.zeros(cop - cout)])
            cout, cin = cop, cip
        self.split = _SPLIT_DEFAULT[0] if split is None else bool(split)
        self.cin, self.cout = cin, cout
        self.kc = pick_kc(cin)
        self.npad = ceil_to(cout, 16)
        self.ntaps = len(mats)
        if self.ntaps > capi.I2R_MAX_TAPS:
            raise ValueError("too many taps")
        self.dy, self.dx = list(dys), list(dxs)
        self.stride, self.relu = stride, relu
        if self.split:
            sc = scale.float()[:cout].reshape(cout, 1)
            self.w = pack_taps(split_virtual_taps(mats), self.kc).to(device)
            folded = split_virtual_taps([m.float() * sc for m in mats])
        else:
            self.w = pack_taps(mats, self.kc).to(device)
            folded = None
        self.scale = pad_vec(scale, self.npad, 1.0).to(device)
        self.bias = pad_vec(bias, self.npad, 0.0).to(device)
        # operand image of the persistent halo kernel (scale folded into the weights, bias block first)
        if folded is not None:
            self.w_folded = pack_folded(folded, torch.ones(cout), bias, self.kc).to(device)
        else:
            self.w_folded = pack_folded(mats, scale, bias, self.kc).to(device)
        self.w_copies = 1
        self._halves = None

    STREAM_COPIES = int(__import__('os').environ.get('I2R_STREAM_COPIES', '1'))

    def replicate(self, copies):
        """Repeat the halo operand image `copies` times (streamed-weight layers: CTA i reads copy i % copies)."""
        if copies > 1 and self.w_copies == 1:
            self.w_folded = self.w_folded.unsqueeze(0).repeat(copies, 1, 1, 1).contiguous()
            self.w_copies = copies

    @property
    def weight_bytes(self):
        return self.w.numel() * 2

    def chunks(self, width=256):
        """Column chunks of at most `width` output channels (the kernels hold <= 256 accumulator columns): list of
        (first channel, sub-layer).  Cached."""
        key = ("chunks", width)
        cache = self.__dict__.setdefault("_chunk_cache", {})
        if key not in cache:
            parts = []
            for c0 in range(0, self.cout, width):
                h = min(width, self.cout - c0)
                sub = ConvLayer.__new__(ConvLayer)
                sub.__dict__.update(self.__dict__)
                sub.cout, sub.npad = h, ceil_to(h, 16)
                assert sub.npad == h, "chunked layers need a multiple of 16 output channels"
                sub.w = self.w[:, :, c0:c0 + h, :].contiguous()
                sub.scale = self.scale[c0:c0 + h].contiguous()
                sub.bias = self.bias[c0:c0 + h].contiguous()
                sub.w_folded = self.w_folded[:, c0:c0 + h, :].contiguous()
                sub._halves = None
                sub._chunk_cache = {}
                parts.append((c0, sub))
            cache[key] = parts
        return cache[key]

    def halves(self):
        """Two layers producing output channels [0, Cout/2) and [Cout/2, Cout) (N-split for the TMA kernel:
        halves the weight footprint per CTA so it can stay resident, doubles the CTA count)."""
        if self._halves is None:
            h = self.cout // 2
            parts = []
            for c0 in (0, h):
                sub = ConvLayer.__new__(ConvLayer)
                sub.__dict__.update(self.__dict__)
                sub.cout, sub.npad = h, h
                sub.w = self.w[:, :, c0:c0 + h, :].contiguous()   # c0 % 8 == 0 keeps the row swizzle phase
                sub.scale = self.scale[c0:c0 + h].contiguous()
                sub.bias = self.bias[c0:c0 + h].contiguous()
                sub.w_folded = self.w_folded[:, c0:c0 + h, :].contiguous()
                sub.replicate(self.STREAM_COPIES)
                sub._halves = None
                parts.append((c0, sub))
            self._halves = parts
        return self._halves


def halo_smem_bytes(p):
    """Shared-memory traffic conv_halo_kernel generates for one problem, in bytes (the kernel's binding resource at
    these widths): per 128-pixel tile the tensor core reads A (128 x 16 fp16 = 4 KB) and B (Npad x 16 fp16)
    for every (tap, K=16 step) MMA plus the bias MMA, the TMA unit writes the halo tile of every 64-channel K-chunk and
    -- when the weight image does not stay resident (> 120 KB) -- the whole weight image, and a staged fp16 output tile
    is written once and read once.  Split-operand problems run 3 K passes and store hi | lo."""
    split = 3 if (p.flags & capi.F_SPLIT) else 1
    npix = p.NB * p.OH * p.OW
    if p.ntaps == 1 and npix % 8 == 0:
        tiles = -(-npix // 128)
    else:
        tiles = p.NB * (-(-p.OW // 8)) * (-(-p.OH // 16))
    ksteps = -(-p.Cin // 16) * split
    mma = (p.ntaps * ksteps + 1) * (128 * 32 + p.Npad * 32)
    halo = 2 if p.ntaps == 9 else 0
    staged = (2 if p.stride == 1 else 3) if split == 3 else 1     # stride-1 split problems stage x_hi / W_hi once
    a_in = -(-p.Cin // 64) * staged * ((8 + halo) * (16 + halo) * 128 if p.stride == 1 else 71808)
    w_image = (p.ntaps * -(-p.Cin // 64) * staged + 1) * p.Npad * 128
    w_in = w_image if w_image > 120 * 1024 else 0
    out = 2 * 128 * p.Cout * 2 * (2 if split == 3 else 1)
    return tiles * (mma + a_in + w_in + out)


def _stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Chain:
    def __init__(self, runner):
        self.r = runner
        self.outer = False

    def __enter__(self):
        r = self.r
        if r._chain is not None or not r.chain_enabled or not r.use_tma or r._in_parallel:
            self.outer = True        # nested, disabled or inside concurrent branches: plain launches
        else:
            r._chain = []
        return self

    def __exit__(self, exc_type, exc, tb):
        r = self.r
        if not self.outer:
            try:
                if exc_type is None:
                    r._flush_chain()
            finally:
                r._chain = None
        return False


class Runner:
    """Launch helper bound to one device.  impl=0: tcgen05 kernels (product); impl=1: scalar check kernel."""

    def __init__(self, device, impl=0):
        self.lib = capi.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise capi.I2RError("the I2R-Net hot path runs on CUDA devices only (got %s)" % device)
        capi.check(self.lib.i2r_device_check(self.device.index or 0), "i2r_device_check")
        self.impl = impl
        self.use_tma = impl == 0   # route eligible problems to the persistent TMA kernel
        self.launches = 0
        self.split = False     # split-operand activations (fp16 hi | lo pairs) for the non-GEMM kernels
        self._stem_images = {}  # packed operand images of the tensor-core stem kernel, keyed by the weight tensors
        self.timing = None     # bench.py: list of (start_event, end_event, algorithmic_flops, nprob) per igemm launch
        self._side_streams = []
        self.concurrent_branches = __import__("os").environ.get("I2R_CONCURRENT_BRANCHES", "1") != "0"
        # chained halo launches (Runner.chain), I2R_HALO_CHAIN=1.  Off by default: bit-identical to separate launches
        # (tests/test_halo_chain_gpu.py) but not faster -- with the hand-shake removed entirely a chained HRNet module
        # takes exactly as long as its eight PDL-overlapped launches (180.6 vs 181.2 us at 32 crops), because each layer's
        # critical path is the CTAs streaming the 192-channel weights, not the launch boundary
        # (profiles/r02_chain_mode.txt)
        self.chain_enabled = __import__("os").environ.get("I2R_HALO_CHAIN", "0") == "1"
        self._chain = None       # open chain: list of layers (lists of ConvProblem)
        self._chain_ws = None    # completion counters of the chained launches (grown on demand, zeroed per launch)
        self._in_parallel = 0
        self.chains = 0          # chained launches made (bench.py / tests)
        self.window_att_tc = __import__("os").environ.get("I2R_WINDOW_ATT", "tc") != "mma"

    # ------------------------------------------------------------------ independent launch chains
    def parallel(self, chains):
        """Run independent launch chains (callables) concurrently: chain 0 on the current stream, the others on side
        streams forked from it and joined back (fork / join by events, so the pattern is capturable into a CUDA graph as
        parallel branches).  Every chain starts after all work queued so far and everything queued afterwards waits for
        all chains, which also keeps the caching allocator's per-stream block reuse safe.  Returns the chains' results.
        The resolution branches of an HRFormer module are such chains: at one image per rank their kernels fill a
        fraction of the SMs each (lib/models/hrformer.py:1714-1731 runs them one after the other)."""
        if len(chains) == 1 or self.device.type != "cuda" or not self.concurrent_branches:
            return [c() for c in chains]
        self._in_parallel += 1      # CTAs of a chained launch wait for each other: never two of them at once
        try:
            return self._parallel(chains)
        finally:
            self._in_parallel -= 1

    def _parallel(self, chains):
        main = torch.cuda.current_stream(self.device)
        while len(self._side_streams) < len(chains) - 1:
            self._side_streams.append(torch.cuda.Stream(device=self.device))
        fork = torch.cuda.Event()
        fork.record(main)
        out = [None] * len(chains)
        joins = []
        for i in range(1, len(chains)):
            st = self._side_streams[i - 1]
            st.wait_event(fork)
            with torch.cuda.stream(st):
                out[i] = chains[i]()
                ev = torch.cuda.Event()
                ev.record(st)
                joins.append(ev)
        out[0] = chains[0]()
        for ev in joins:
            main.wait_event(ev)
        return out

    # ------------------------------------------------------------------ implicit GEMM
    def problem(self, L, x, out=None, add0=None, add0_shift=0, add1=None, add1_shift=0, in_shift=0,
                relu=None, out_mode="nhwc16", out_hw=None, out_mul=1, out_off=(0, 0), ohow=None, gelu=False,
                act_first=False, lo_offset=0):
        """Build one problem.  x: fp16 [NB, Hs, Ws, C>=Cin] (channel stride 1).  Returns (ConvProblem, out).
        lo_offset > 0 (split-operand mode only): `out` / `add0` / `add1` are the hi-half channel SLICES [.., Cout] of
        wider pair tensors whose lo halves sit lo_offset channels further (one layer run as several problems)."""
        assert x.dtype == torch.float16 and x.dim() == 4 and x.stride(3) == 1
        nb, hs, ws, cphys = x.shape
        assert cphys >= (2 * L.cin if L.split else L.cin), "input tensor has too few channels for this layer"
        cw = 2 * L.cout if L.split else L.cout     # channels of an fp16 output / addend tensor (split: hi | lo)
        pix = x.stride(2)
        assert x.stride(1) == ws * pix and (nb == 1 or x.stride(0) == hs * ws * pix), "x must be pixel-contiguous"
        ih, iw = hs << in_shift, ws << in_shift
        if ohow is None:
            if L.stride == 1:
                oh, ow = ih, iw
            else:
                oh, ow = (ih + L.stride - 1) // L.stride, (iw + L.stride - 1) // L.stride
        else:
            oh, ow = ohow
        ohf, owf = out_hw if out_hw is not None else (oh * out_mul, ow * out_mul)
        flags = 0
        if (L.relu if relu is None else relu):
            flags |= capi.F_RELU
        if gelu:                 # erf-GELU instead of ReLU
            flags = (flags & ~capi.F_RELU) | capi.F_GELU
        if act_first:            # y = act(conv) + addends
            flags |= capi.F_ACT_FIRST
        if out is None:
            if out_mode == "nhwc16":
                out = torch.empty((nb, ohf, owf, cw), dtype=torch.float16, device=x.device)
            elif out_mode == "nchw32":
                out = torch.empty((nb, L.cout, ohf, owf), dtype=torch.float32, device=x.device)
            elif out_mode == "nhwc32":
                out = torch.empty((nb, ohf, owf, L.cout), dtype=torch.float32, device=x.device)
            elif out_mode == "t16":      # channel-major [cw, pixels] (halo kernel only): V^T for the tcgen05 attention
                out = torch.empty((cw, nb * ohf * owf), dtype=torch.float16, device=x.device)
            else:
                raise ValueError(out_mode)
        if out_mode == "nchw32":
            flags |= capi.F_OUT_NCHW_F32
            out_pix = 0
            assert out.is_contiguous()
        elif out_mode == "t16":
            flags |= capi.F_OUT_T16
            assert out.is_contiguous() and out.shape[1] == nb * ohf * owf
            out_pix = out.stride(0)
        else:
            if out_mode == "nhwc32":
                flags |= capi.F_OUT_F32
            assert out.stride(3) == 1
            out_pix = out.stride(2)
        p = capi.ConvProblem()
        p.x, p.w, p.scale, p.bias = x.data_ptr(), L.w.data_ptr(), L.scale.data_ptr(), L.bias.data_ptr()
        p.add0 = add0.data_ptr() if add0 is not None else None
        p.add1 = add1.data_ptr() if add1 is not None else None
        add_pix = cw
        for a in (add0, add1):
            if a is not None:
                assert a.dtype == torch.float16 and a.dim() == 4 and a.stride(3) == 1
                assert a.shape[-1] == (L.cout if lo_offset else cw)
                assert a.stride(1) == a.shape[2] * a.stride(2)
                add_pix = a.stride(2)
        if add0 is not None and add1 is not None:
            assert add0.stride(2) == add1.stride(2), "both addends must share one pixel stride"
        p.y = out.data_ptr()
        p.NB, p.IH, p.IW, p.Cin, p.KC = nb, ih, iw, L.cin, L.kc
        p.in_pix_stride, p.in_shift = pix, in_shift
        p.OH, p.OW, p.stride = oh, ow, L.stride
        p.Cout, p.Npad, p.out_pix_stride = L.cout, L.npad, out_pix
        p.OHf, p.OWf = ohf, owf
        p.out_mul, p.out_offy, p.out_offx = out_mul, out_off[0], out_off[1]
        p.add0_shift, p.add1_shift = add0_shift, add1_shift
        p.add_pix_stride = add_pix
        p.ntaps = L.ntaps
        for t in range(L.ntaps):
            p.dy[t], p.dx[t] = L.dy[t], L.dx[t]
        if L.split:
            flags |= capi.F_SPLIT
        p.flags = flags
        p.w_folded = L.w_folded.data_ptr()
        p.w_folded_copies = L.w_copies
        p.pair_lo_offset = lo_offset
        assert not lo_offset or (L.split and out_mode == "nhwc16")
        p._keep = (x, L, add0, add1, out)
        return p, out

    TMA_WEIGHT_RESIDENT_MAX = 120 * 1024
    # N-split of layers whose weight image exceeds shared memory into two half-width problems (default).  With
    # I2R_HALO_PAIR=1/2 such layers run as ONE problem on CTA pairs instead (tcgen05 cta_group::2: each CTA of a pair holds
    # half of the weight rows, nobody re-reads the activations; csrc/conv_halo.cu explains why that is not the default).
    N_SPLIT = __import__("os").environ.get("I2R_HALO_PAIR", "0") == "0"

    def problems(self, L, x, **kw):
        """Like problem(), but may split the output channels in two (see ConvLayer.halves).  Returns (list, out)."""
        std = (L.ntaps == 1 and L.dy[0] == 0 and L.dx[0] == 0) or (
            L.ntaps == 9 and all(L.dy[t] == t // 3 - 1 and L.dx[t] == t % 3 - 1 for t in range(9)))
        # stride-2 3x3 problems (parity-plane staging, two 71 KB activation stages) only have room for weight-ring slots
        # of <= 96 output channels: wider layers run as two halves as well
        s2_wide = L.stride == 2 and L.ntaps == 9 and L.npad > 96
        split = (self.use_tma and std and (L.stride == 1 or s2_wide) and kw.get("in_shift", 0) == 0 and
                 kw.get("out_mul", 1) == 1 and kw.get("out_mode", "nhwc16") == "nhwc16" and
                 kw.get("add0_shift", 0) == 0 and kw.get("add1_shift", 0) == 0 and
                 L.weight_bytes > self.TMA_WEIGHT_RESIDENT_MAX and L.cout == L.npad and L.cout % 32 == 0 and
                 not L.split and self.N_SPLIT)
        if L.npad > 256:
            # more output channels than accumulator columns: column chunks writing channel slices of one tensor
            assert kw.get("out_mode", "nhwc16") == "nhwc16" and L.cout == L.npad
            out = kw.pop("out", None)
            if out is None:
                if kw.get("ohow"):
                    oh, ow = kw["ohow"]
                else:
                    st, sh = L.stride, kw.get("in_shift", 0)
                    oh, ow = ((x.shape[1] << sh) + st - 1) // st, ((x.shape[2] << sh) + st - 1) // st
                out = torch.empty((x.shape[0], oh, ow, (2 if L.split else 1) * L.cout), dtype=torch.float16,
                                  device=x.device)
            add0, add1 = kw.pop("add0", None), kw.pop("add1", None)
            probs = []
            for c0, sub in L.chunks(256):
                sl = slice(c0, c0 + sub.cout)
                p, _ = self.problem(sub, x, out=out[..., sl], add0=None if add0 is None else add0[..., sl],
                                    add1=None if add1 is None else add1[..., sl],
                                    lo_offset=L.cout if L.split else 0, **kw)
                probs.append(p)
            return probs, out
        if not split:
            p, out = self.problem(L, x, **kw)
            return [p], out
        out = kw.pop("out", None)
        if out is None:
            oh, ow = kw["ohow"] if kw.get("ohow") else ((x.shape[1] + L.stride - 1) // L.stride,
                                                        (x.shape[2] + L.stride - 1) // L.stride)
            out = torch.empty((x.shape[0], oh, ow, L.cout), dtype=torch.float16, device=x.device)
        add0, add1 = kw.pop("add0", None), kw.pop("add1", None)
        probs = []
        for c0, sub in L.halves():
            sl = slice(c0, c0 + sub.cout)
            p, _ = self.problem(sub, x, out=out[..., sl], add0=None if add0 is None else add0[..., sl],
                                add1=None if add1 is None else add1[..., sl], **kw)
            probs.append(p)
        return probs, out

    # ------------------------------------------------------------------ chained launches
    def chain(self):
        """Context manager: the conv / conv_group launches issued inside -- dependent layers of halo-kernel problems,
        e.g. the conv1 / conv2 sequence of an HRNet module's BasicBlocks -- are collected and run as ONE persistent grid
        (i2r_conv_halo_chain: per-image completion counters instead of launch boundaries).  Outputs are allocated as
        usual and must not be read by anything else before the context closes.  Falls back to one launch per layer when
        the chain does not qualify; a launch with problems the halo kernel does not take flushes the chain first."""
        return _Chain(self)

    def _flush_chain(self):
        layers, self._chain = self._chain, []
        if not layers:
            return
        chained = False
        if len(layers) > 1:
            flat = [p for layer in layers for p in layer]
            arr = (capi.ConvProblem * len(flat))(*flat)
            need = int(self.lib.i2r_conv_halo_chain_workspace(arr, len(flat)))
            if self._chain_ws is None or self._chain_ws.numel() * 4 < need:
                self._chain_ws = torch.zeros(max(4096, need // 4 * 2), dtype=torch.int32, device=self.device)
            counts = (ctypes.c_int * len(layers))(*[len(layer) for layer in layers])
            if self.timing is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            rc = self.lib.i2r_conv_halo_chain(arr, counts, len(layers), self._chain_ws.data_ptr(),
                                              self._chain_ws.numel() * 4, _stream_ptr())
            if rc == 0:
                chained = True
                self.launches += 1
                self.chains += 1
                if self.timing is not None:
                    e1.record()
                    flops = sum(2.0 * p.NB * p.OH * p.OW * p.Cout * p.Cin * p.ntaps for p in flat)
                    self.timing.append((e0, e1, flops, len(flat), sum(halo_smem_bytes(p) for p in flat)))
            elif rc != capi.E_UNSUPPORTED:
                capi.check(rc, "i2r_conv_halo_chain")
        if not chained:
            for layer in layers:
                self._launch_now(layer)

    def launch(self, problems):
        if self._chain is not None and not self._in_parallel:
            halo = all(not (p.flags & (capi.F_OUT_T16 | capi.F_OUT_F32 | capi.F_OUT_NCHW_F32)) and self.use_tma and
                       p.stride == 1 and
                       self.lib.i2r_conv_halo_supported(ctypes.byref(p)) for p in problems)
            room = (len(self._chain) < capi.I2R_MAX_CHAIN_LAYERS and len(problems) <= capi.I2R_MAX_GROUP and
                    sum(len(layer) for layer in self._chain) + len(problems) <= capi.I2R_MAX_CHAIN_PROBLEMS)
            if halo and not room:
                self._flush_chain()
            if halo:
                self._chain.append(list(problems))
                return
            self._flush_chain()
        self._launch_now(problems)

    def _launch_now(self, problems):
        if self.timing is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        tma, gen = [], []
        for p in problems:
            t16 = bool(p.flags & capi.F_OUT_T16)      # transposed output exists in the halo kernel only
            halo = (self.use_tma or t16) and self.lib.i2r_conv_halo_supported(ctypes.byref(p))
            if t16 and not halo:
                raise capi.I2RError("transposed (t16) output needs a problem the halo kernel supports")
            (tma if halo else gen).append(p)
        for group, fn in ((tma, "tma"), (gen, "igemm")):
            for i in range(0, len(group), capi.I2R_MAX_GROUP):
                chunk = group[i:i + capi.I2R_MAX_GROUP]
                arr = (capi.ConvProblem * len(chunk))(*chunk)
                if fn == "tma":
                    capi.check(self.lib.i2r_conv_halo(arr, len(chunk), _stream_ptr()), "i2r_conv_halo")
                else:
                    capi.check(self.lib.i2r_conv_igemm(arr, len(chunk), self.impl, _stream_ptr()), "i2r_conv_igemm")
                self.launches += 1
        if self.timing is not None:
            e1.record()
            flops = sum(2.0 * p.NB * p.OH * p.OW * p.Cout * p.Cin * p.ntaps for p in problems)
            smem = sum(halo_smem_bytes(p) for p in tma) if not gen else 0
            self.timing.append((e0, e1, flops, len(problems), smem))

    def conv(self, L, x, **kw):
        probs, out = self.problems(L, x, **kw)
        self.launch(probs)
        return out

    def conv_group(self, specs):
        """specs: list of (L, x, kwargs) launched together (one grid per kernel kind).  Returns the outputs."""
        probs, outs = [], []
        for L, x, kw in specs:
            ps, o = self.problems(L, x, **dict(kw))
            probs.extend(ps)
            outs.append(o)
        self.launch(probs)
        return outs

    def linear(self, L, x2d, add0=None, relu=None):
        """x2d: fp16 [T, Cin] (row stride >= Cin) -> contiguous [T, Cout]: a 1x1 'convolution' over T pixels."""
        p, out = self.linear_problem(L, x2d, add0=add0, relu=relu)
        self.launch([p])
        return out.view(x2d.shape[0], -1)

    def linear_problem(self, L, x2d, add0=None, relu=None, out_mode="nhwc16"):
        t, c = x2d.shape
        assert x2d.stride(1) == 1
        ld = x2d.stride(0)
        x4 = x2d.as_strided((1, t, 1, c), (t * ld, ld, ld, 1))
        a4 = add0.as_strided((1, t, 1, add0.shape[1]), (t * add0.stride(0), add0.stride(0), add0.stride(0), 1)) \
            if add0 is not None else None
        return self.problem(L, x4, add0=a4, relu=relu, out_mode=out_mode)

    # ------------------------------------------------------------------ small kernels
    def stem(self, x, w, scale, bias, cout):
        nb, cin, h, wd = x.shape
        assert x.dtype == torch.float32 and x.is_contiguous()
        y = torch.empty((nb, h // 2, wd // 2, cout * (2 if self.split else 1)), dtype=torch.float16, device=x.device)
        if self.impl == 0:     # tensor-core kernel (product); the SIMT kernel below stays as the check implementation
            key = (id(w), id(scale))        # the entry keeps w / scale alive, so the ids cannot be recycled
            ent = self._stem_images.get(key)
            if ent is None:
                from .packing import pack_stem_tc
                ent = (pack_stem_tc(w.detach().cpu(), scale.detach().cpu()).to(x.device), w, scale)
                self._stem_images[key] = ent
            img = ent[0]
            capi.check(self.lib.i2r_stem_conv3x3s2_tc(x.data_ptr(), img.data_ptr(), bias.data_ptr(), y.data_ptr(), nb,
                                                       cin, h, wd, cout, int(self.split), _stream_ptr()),
                       "i2r_stem_conv3x3s2_tc")
            self.launches += 1
            return y
        capi.check(self.lib.i2r_stem_conv3x3s2(x.data_ptr(), w.data_ptr(), scale.data_ptr(), bias.data_ptr(),
                                                y.data_ptr(), nb, cin, h, wd, cout, int(self.split), _stream_ptr()),
                   "i2r_stem_conv3x3s2")
        self.launches += 1
        return y

    def mask_res_stem(self, mask, w_pre, w1, scale, bias):
        """conv_pre 3x3 (1 -> 3) + resnet18 conv1 7x7 s2 (3 -> 64) + bn1 + ReLU on fp32 masks [S,1,H,W] -> fp16 NHWC."""
        nb, cin, h, wd = mask.shape
        assert cin == 1 and mask.dtype == torch.float32 and mask.is_contiguous()
        y = torch.empty((nb, h // 2, wd // 2, 64 * (2 if self.split else 1)), dtype=torch.float16, device=mask.device)
        capi.check(self.lib.i2r_mask_res_stem(mask.data_ptr(), w_pre.data_ptr(), w1.data_ptr(), scale.data_ptr(),
                                              bias.data_ptr(), y.data_ptr(), nb, h, wd, int(self.split), _stream_ptr()),
                   "i2r_mask_res_stem")
        self.launches += 1
        return y

    def maxpool(self, x):
        nb, h, w, c = x.shape
        assert x.is_contiguous() and x.dtype == torch.float16
        y = torch.empty((nb, (h + 1) // 2, (w + 1) // 2, c), dtype=torch.float16, device=x.device)
        capi.check(self.lib.i2r_maxpool3x3s2(x.data_ptr(), y.data_ptr(), nb, h, w, c // 2 if self.split else c,
                                              int(self.split), _stream_ptr()),
                   "i2r_maxpool3x3s2")
        self.launches += 1
        return y

    def layernorm(self, x2d, gamma, beta, eps=1e-5, pos=None):
        rows, c = x2d.shape
        assert x2d.is_contiguous() and x2d.dtype == torch.float16
        y = torch.empty_like(x2d)
        y2 = torch.empty_like(x2d) if pos is not None else None
        capi.check(self.lib.i2r_layernorm(x2d.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                           pos.data_ptr() if pos is not None else None, y.data_ptr(),
                                           y2.data_ptr() if y2 is not None else None, rows,
                                           c // 2 if self.split else c, eps, int(self.split), _stream_ptr()),
                   "i2r_layernorm")
        self.launches += 1
        return y, y2

    def add(self, a, b):
        assert a.shape == b.shape and a.is_contiguous() and b.is_contiguous()
        y = torch.empty_like(a)
        capi.check(self.lib.i2r_add_f16(a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(),
                                         a.shape[-1] // 2 if self.split else 0, _stream_ptr()),
                   "i2r_add_f16")
        self.launches += 1
        return y

    def upsum(self, x0, t1, shift1, t2=None, shift2=0, relu=True):
        """relu(x0 + up(t1) + up(t2)) with nearest upsampling by 2^shift (HRNet fuse, highest-resolution branch)."""
        nb, h, w, c = x0.shape
        assert x0.is_contiguous() and t1.is_contiguous() and (t2 is None or t2.is_contiguous())
        assert t1.shape == (nb, h >> shift1, w >> shift1, c) and (t2 is None or t2.shape == (nb, h >> shift2, w >> shift2, c))
        y = torch.empty_like(x0)
        capi.check(self.lib.i2r_upsum(x0.data_ptr(), t1.data_ptr(), shift1, t2.data_ptr() if t2 is not None else None,
                                      shift2, y.data_ptr(), nb, h, w, c // 2 if self.split else c, int(relu),
                                      int(self.split), _stream_ptr()), "i2r_upsum")
        self.launches += 1
        return y

    def attention(self, q, k, v, cu_seqlens, max_seqlen, scale, lo=None):
        """q,k,v: fp16 2-D views [T, D] (row stride arbitrary multiple of 8); returns [T, D] contiguous.
        Split-operand mode: `lo` = (q_lo, k_lo, v_lo) element offsets from each hi view to its lo half; the result is
        a split tensor [T, 2D]."""
        t, d = q.shape
        out = torch.empty((t, 2 * d if lo is not None else d), dtype=torch.float16, device=q.device)
        nseq = cu_seqlens.numel() - 1
        ws_bytes = int(self.lib.i2r_attention_workspace_bytes(t, d, nseq, max_seqlen))
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=q.device) if ws_bytes else None
        ql, kl, vl = lo if lo is not None else (0, 0, 0)
        capi.check(self.lib.i2r_attention_varlen(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(),
                                                  q.stride(0), k.stride(0), v.stride(0), out.stride(0), d,
                                                  cu_seqlens.data_ptr(), nseq, max_seqlen, t, scale,
                                                  ws.data_ptr() if ws is not None else None, ws_bytes,
                                                  int(lo is not None), ql, kl, vl, d, _stream_ptr()),
                   "i2r_attention_varlen")
        self.launches += 1
        return out

    def attention_tc(self, q, k, vt, cu_seqlens, max_seqlen, scale, split=False):
        """tcgen05 attention.  q, k: fp16 2-D views [T, D] (split: [T, 2D] pairs, lo directly after hi), row strides
        multiples of 8; vt: V transposed [D or 2D, T] (Runner.problem(out_mode="t16")).  Returns [T, D] ([T, 2D])."""
        t, w = q.shape
        d = w // 2 if split else w
        assert k.shape == q.shape and vt.shape == (w, t) and vt.stride(1) == 1 and q.stride(1) == 1 and k.stride(1) == 1
        out = torch.empty((t, w), dtype=torch.float16, device=q.device)
        nseq = cu_seqlens.numel() - 1
        ws_bytes = int(self.lib.i2r_attention_tc_workspace_bytes(t, d, nseq, max_seqlen))
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=q.device) if ws_bytes else None
        capi.check(self.lib.i2r_attention_tc(q.data_ptr(), k.data_ptr(), vt.data_ptr(), out.data_ptr(), q.stride(0),
                                              k.stride(0), vt.stride(0), out.stride(0), d, cu_seqlens.data_ptr(), nseq,
                                              max_seqlen, t, scale, ws.data_ptr() if ws is not None else None,
                                              ws_bytes, int(split), d, _stream_ptr()), "i2r_attention_tc")
        self.launches += 1 + (1 if ws_bytes else 0)
        return out

    def encoder_tail(self, tail, attn, src, pos=None, eps=1e-5):
        """Fused out-proj + residual + LN1 + FFN + residual + LN2 (+ pos).  tail: EncoderTailParams; attn / src / pos:
        fp16 [T, 96] ([T, 192] pairs in split mode).  Returns (src_next, src_next + pos or None)."""
        t, w = attn.shape
        assert src.shape == (t, w) and attn.stride(1) == 1 and src.stride(1) == 1 and w == (192 if tail.split else 96)
        out = torch.empty((t, w), dtype=torch.float16, device=attn.device)
        out_pos = torch.empty_like(out) if pos is not None else None
        if pos is not None:
            assert pos.shape == (t, w) and pos.is_contiguous()
        capi.check(self.lib.i2r_encoder_tail(attn.data_ptr(), attn.stride(0), src.data_ptr(), src.stride(0),
                                              pos.data_ptr() if pos is not None else None, out.data_ptr(),
                                              out_pos.data_ptr() if out_pos is not None else None, w,
                                              tail.wimg.data_ptr(), tail.params.data_ptr(), t, 96, 192, eps,
                                              int(tail.split), _stream_ptr()), "i2r_encoder_tail")
        self.launches += 1
        return out, out_pos

    # ------------------------------------------------------------------ HRFormer-B building blocks
    ACT = {None: 0, "none": 0, "relu": 1, "gelu": 2}

    def dwconv3x3(self, x, w, scale, bias, stride=1, act=None):
        """Depthwise 3x3 (pad 1) + scale/bias + activation.  x: fp16 NHWC (pair tensor when self.split); w fp32 [9, C]."""
        nb, h, wd, cw = x.shape
        c = cw // 2 if self.split else cw
        assert x.is_contiguous() and w.shape == (9, c) and w.dtype == torch.float32 and w.is_contiguous()
        y = torch.empty((nb, (h + stride - 1) // stride, (wd + stride - 1) // stride, cw), dtype=torch.float16,
                        device=x.device)
        capi.check(self.lib.i2r_dwconv3x3(x.data_ptr(), w.data_ptr(), scale.data_ptr(), bias.data_ptr(), y.data_ptr(),
                                           nb, h, wd, c, stride, self.ACT[act], int(self.split), _stream_ptr()),
                   "i2r_dwconv3x3")
        self.launches += 1
        return y

    def upsum_bilinear(self, x0, terms, relu=True):
        """relu(x0 + sum bilinear_up(t, 2^shift)) (align_corners=False).  terms: up to three (tensor, shift)."""
        nb, h, w, cw = x0.shape
        c = cw // 2 if self.split else cw
        assert 1 <= len(terms) <= 3 and x0.is_contiguous()
        args = []
        for t, s in list(terms) + [(None, 0)] * (3 - len(terms)):
            assert t is None or (t.is_contiguous() and t.shape == (nb, h >> s, w >> s, cw))
            args += [t.data_ptr() if t is not None else None, s]
        y = torch.empty_like(x0)
        capi.check(self.lib.i2r_upsum_bilinear(x0.data_ptr(), *args, y.data_ptr(), nb, h, w, c, int(relu),
                                                int(self.split), _stream_ptr()), "i2r_upsum_bilinear")
        self.launches += 1
        return y

    def layernorm_padded(self, x2d, gamma, beta, c_real, eps=1e-6):
        """LayerNorm over the first c_real channels of rows padded to C_pad channels (pad channels come out zero)."""
        rows, cw = x2d.shape
        cp = cw // 2 if self.split else cw
        assert x2d.is_contiguous() and gamma.numel() == cp and beta.numel() == cp
        y = torch.empty_like(x2d)
        capi.check(self.lib.i2r_layernorm_padded(x2d.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), rows,
                                                  c_real, cp, eps, int(self.split), _stream_ptr()),
                   "i2r_layernorm_padded")
        self.launches += 1
        return y

    def ln_window_gather(self, x, gamma, beta, c_real, ws=7, eps=1e-6):
        """LayerNorm + window-major (padded) token layout.  x: [NB, H, W, C_pad] -> [NB*Hp*Wp, C_pad] (pair when split)."""
        nb, h, w, cw = x.shape
        cp = cw // 2 if self.split else cw
        assert x.is_contiguous() and gamma.numel() == cp
        rows = int(self.lib.i2r_window_rows(nb, h, w, ws))
        y = torch.empty((rows, cw), dtype=torch.float16, device=x.device)
        capi.check(self.lib.i2r_ln_window_gather(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), nb, h, w,
                                                  c_real, cp, ws, eps, int(self.split), _stream_ptr()),
                   "i2r_ln_window_gather")
        self.launches += 1
        return y

    def window_scatter_add(self, x, a, ws=7):
        """x + (window-major rows `a` scattered back to pixels, padded positions dropped)."""
        nb, h, w, cw = x.shape
        c = cw // 2 if self.split else cw
        assert x.is_contiguous() and a.is_contiguous() and a.shape == (int(self.lib.i2r_window_rows(nb, h, w, ws)), cw)
        y = torch.empty_like(x)
        capi.check(self.lib.i2r_window_scatter_add(x.data_ptr(), a.data_ptr(), y.data_ptr(), nb, h, w, c, ws,
                                                    int(self.split), _stream_ptr()), "i2r_window_scatter_add")
        self.launches += 1
        return y

    def window_attention(self, q, k, v, win_len, heads, scale, head_pad=48):
        """q, k, v: fp16 2-D views [T, heads*head_pad] (split: hi views of pair rows [.., 2*heads*head_pad]); T = nwin *
        win_len.  Returns [T, heads*head_pad] ([T, 2*heads*head_pad] pair)."""
        t, cq = q.shape
        assert cq == heads * head_pad and t % win_len == 0 and q.stride(1) == 1
        out = torch.empty((t, 2 * cq if self.split else cq), dtype=torch.float16, device=q.device)
        lo = cq if self.split else 0
        # product path: tcgen05 kernel (49-token windows, 48-channel heads); impl=1 / I2R_WINDOW_ATT=mma: mma.sync kernel
        fn = self.lib.i2r_window_attention
        if self.impl == 0 and win_len == 49 and head_pad == 48 and self.window_att_tc:
            fn = self.lib.i2r_window_attention_tc
        capi.check(fn(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), q.stride(0),
                                                  k.stride(0), v.stride(0), out.stride(0), t // win_len, win_len, heads,
                                                  head_pad, scale, int(self.split), lo, lo, lo, lo, _stream_ptr()),
                   "i2r_window_attention")
        self.launches += 1
        return out



def _flushes_chain(fn):
    def wrapped(self, *a, **kw):
        if self._chain:          # an open chain with deferred layers: they come first in stream order
            self._flush_chain()
        return fn(self, *a, **kw)
    wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
    return wrapped


for _name in ("parallel", "stem", "mask_res_stem", "maxpool", "layernorm", "add", "upsum", "attention", "attention_tc",
              "encoder_tail", "dwconv3x3", "upsum_bilinear", "layernorm_padded", "ln_window_gather", "window_scatter_add",
              "window_attention"):
    setattr(Runner, _name, _flushes_chain(getattr(Runner, _name)))

class EncoderTailParams:
    """Device-resident operand image of i2r_encoder_tail for one encoder layer."""

    def __init__(self, w_out, b_out, w1, b1, w2, b2, g1, be1, g2, be2, device="cuda", split=None):
        from .packing import pack_encoder_tail
        self.split = _SPLIT_DEFAULT[0] if split is None else bool(split)
        self.host = tuple(t.detach().float().cpu() for t in (w_out, b_out, w1, b1, w2, b2, g1, be1, g2, be2))
        img, params = pack_encoder_tail(*self.host, split=self.split)
        self.wimg = img.to(device)
        self.params = params.to(device)
