"""BatchNorm folding and weight packing into the shared-memory operand layout of the tcgen05 kernel.

Packed B operand (see include/i2r.h, i2r_conv_problem::w):  fp16 [ntaps][Cin/KC][Npad][64]
-- for every tap and K-chunk, Npad rows (output channels) of 128 bytes (KC real K-slots + zero pad),
16-byte chunks XOR-swizzled by the row index: exactly the SWIZZLE_128B K-major image the kernels
bulk-copy into shared memory, so no device-side reordering exists.
"""
import torch

BN_EPS = 1e-5


def ceil_to(x, m):
    return (x + m - 1) // m * m


KC = 64   # K-chunk = 64 channel slots = one 128-byte SWIZZLE_128B row


def pick_kc(cin):
    """All problems use 64-slot K-chunks; the last chunk may be partially filled (Cin multiple of 16)."""
    if cin % 16 != 0:
        raise ValueError("Cin=%d is not a multiple of 16 (channel padding not implemented)" % cin)
    return KC


def fold_bn(bn_sd, prefix, cout, conv_bias=None, eps=BN_EPS):
    """scale/bias (fp32 [cout]) of eval-mode BatchNorm2d `prefix` applied after a conv (+bias)."""
    gamma = bn_sd[prefix + ".weight"].float()
    beta = bn_sd[prefix + ".bias"].float()
    mean = bn_sd[prefix + ".running_mean"].float()
    var = bn_sd[prefix + ".running_var"].float()
    scale = gamma / torch.sqrt(var + eps)
    bias = beta - mean * scale
    if conv_bias is not None:
        bias = bias + conv_bias.float() * scale
    assert scale.numel() == cout
    return scale, bias


def pack_taps(mats, kc=KC):
    """mats: list over taps of fp32 [Cout, Cin] matrices -> fp16 [ntaps, ceil(Cin/64), Npad, 64].

    Per (tap, K-chunk) block: Npad rows (output channels) of 128 B = 64 fp16 K-slots holding input
    channels [64*chunk, 64*chunk+64) (zero past Cin); the eight 16-byte chunks of row n are stored at
    chunk position (c XOR (n & 7)) -- the SWIZZLE_128B K-major image the kernels bulk-copy into shared memory.
    """
    assert kc == KC
    cout, cin = mats[0].shape
    npad = ceil_to(cout, 16)
    nch = (cin + KC - 1) // KC
    out = torch.zeros(len(mats), nch, npad, 8, 8, dtype=torch.float16)
    rows = torch.arange(cout)
    for t, m in enumerate(mats):
        mm = torch.zeros(cout, nch * KC)
        mm[:, :cin] = m.float()
        mm = mm.reshape(cout, nch, 8, 8).to(torch.float16)
        for c in range(8):
            out[t, :, rows, c ^ (rows & 7), :] = mm[:, :, c, :].permute(1, 0, 2)
    return out.reshape(len(mats), nch, npad, 64).contiguous()


def unpack_taps(packed, cin):
    """Inverse of pack_taps: fp16 [ntaps, nch, Npad, 64] -> fp32 [ntaps, Npad, cin] (tests / emulator)."""
    ntaps, nch, npad, _ = packed.shape
    p5 = packed.reshape(ntaps, nch, npad, 8, 8).float()
    rows = torch.arange(npad)
    out = torch.zeros(ntaps, nch, npad, 8, 8)
    for c in range(8):
        out[:, :, rows, c, :] = p5[:, :, rows, c ^ (rows & 7), :]
    return out.permute(0, 2, 1, 3, 4).reshape(ntaps, npad, nch * KC)[:, :, :cin]


def conv_taps(weight, pad):
    """Conv2d weight [Cout,Cin,KH,KW] -> (list of [Cout,Cin] per tap, dy list, dx list)."""
    cout, cin, kh, kw = weight.shape
    mats, dys, dxs = [], [], []
    for ky in range(kh):
        for kx in range(kw):
            mats.append(weight[:, :, ky, kx])
            dys.append(ky - pad)
            dxs.append(kx - pad)
    return mats, dys, dxs


def deconv4x4s2_phase_taps(weight, py, px):
    """ConvTranspose2d(k=4, s=2, p=1) weight [Cin,Cout,4,4]: taps of output phase (oy%2, ox%2) = (py, px).

    out[2q+py] = sum over (ky, dy): py=0 -> (1, 0), (3, -1);  py=1 -> (0, +1), (2, 0)   (same for x).
    """
    sel = {0: ((1, 0), (3, -1)), 1: ((0, 1), (2, 0))}
    mats, dys, dxs = [], [], []
    for ky, dy in sel[py]:
        for kx, dx in sel[px]:
            mats.append(weight[:, :, ky, kx].t())
            dys.append(dy)
            dxs.append(dx)
    return mats, dys, dxs


def pad_vec(v, npad, fill=0.0):
    out = torch.full((npad,), fill, dtype=torch.float32)
    out[: v.numel()] = v.float()
    return out


def pack_folded(mats, scale, bias, kc=KC):
    """Operand image of the persistent halo kernel: fp16 [1 + ntaps*nch, Npad, 64].

    Block 0 is the bias block -- row n holds fp16(bias[n]) in K slot 0 and fp16(bias[n] - slot0) in slot 1, so one
    K=16 MMA against a tile of ones initialises the accumulator with the bias at ~22 bits; blocks 1.. are
    pack_taps() of the weights with the BatchNorm scale folded in (fp32 product, one fp16 rounding)."""
    cout, _ = mats[0].shape
    npad = ceil_to(cout, 16)
    sc = scale.float()[:cout].reshape(cout, 1)
    taps = pack_taps([m.float() * sc for m in mats], kc)
    ntaps, nch = taps.shape[0], taps.shape[1]
    b = bias.float()[:cout]
    hi = b.to(torch.float16)
    lo = (b - hi.float()).to(torch.float16)
    blk = torch.zeros(npad, 8, 8, dtype=torch.float16)
    rows = torch.arange(cout)
    blk[rows, 0 ^ (rows & 7), 0] = hi
    blk[rows, 0 ^ (rows & 7), 1] = lo
    return torch.cat([blk.reshape(1, npad, 64), taps.reshape(ntaps * nch, npad, 64)], 0).contiguous()


def split_virtual_taps(mats):
    """Split-operand mode: activations travel as fp16 pairs (hi | lo, value = hi + lo) and a product is evaluated as
    x_hi*W_hi + x_lo*W_hi + x_hi*W_lo (fp32 accumulation; the dropped lo*lo term is ~2^-22 relative).  Along K that
    is ONE longer GEMM: A chunks [x_hi | x_lo | x_hi] against B chunks [W_hi | W_hi | W_lo].  Returns the per-tap
    virtual weight matrices [Cout, 3 * ceil(Cin/64) * 64] (each third zero-padded to whole 64-slot K-chunks)."""
    out = []
    for m in mats:
        m = m.float()
        cout, cin = m.shape
        kp = ceil_to(cin, KC)
        hi = m.to(torch.float16).float()
        lo = (m - hi).to(torch.float16).float()
        v = torch.zeros(cout, 3 * kp)
        v[:, :cin] = hi
        v[:, kp:kp + cin] = hi
        v[:, 2 * kp:2 * kp + cin] = lo
        out.append(v)
    return out


def split_pair(t):
    """fp32 tensor [..., C] -> fp16 [..., 2C] = (hi | lo)."""
    hi = t.float().to(torch.float16)
    lo = (t.float() - hi.float()).to(torch.float16)
    return torch.cat([hi, lo], dim=-1)


def merge_pair(t):
    """fp16 [..., 2C] (hi | lo) -> fp32 [..., C]."""
    c = t.shape[-1] // 2
    return t[..., :c].float() + t[..., c:].float()


def _pack_rows_sw128(m):
    """fp32 [N, K] (N % 8 == 0) -> fp16 [ceil(K/64), N, 64]: K-major SWIZZLE_128B chunks (16-byte groups XOR row & 7)."""
    n, k = m.shape
    nch = (k + KC - 1) // KC
    mm = torch.zeros(n, nch * KC)
    mm[:, :k] = m.float()
    mm = mm.reshape(n, nch, 8, 8).to(torch.float16)
    out = torch.zeros(nch, n, 8, 8, dtype=torch.float16)
    rows = torch.arange(n)
    for c in range(8):
        out[:, rows, c ^ (rows & 7), :] = mm[:, :, c, :].permute(1, 0, 2)
    return out.reshape(nch, n, 64)


def pack_encoder_tail(w_out, b_out, w1, b1, w2, b2, g1, be1, g2, be2, split):
    """Operand image of i2r_encoder_tail (csrc/encoder_tail.cu): five [96 x 96] matrices W_o, W_1[0:96], W_1[96:192],
    W_2[:, 0:96], W_2[:, 96:192] as SWIZZLE_128B K-major chunks (split: K = [W_hi | W_lo]) and the fp32 vector
    [b_o | b_1 | b_2 | gamma1 | beta1 | gamma2 | beta2]."""
    d = w_out.shape[0]
    assert w_out.shape == (d, d) and w1.shape == (2 * d, d) and w2.shape == (d, 2 * d) and d == 96
    mats = [w_out, w1[:d], w1[d:], w2[:, :d], w2[:, d:]]
    blocks = []
    for m in mats:
        m = m.float()
        if split:
            hi = m.to(torch.float16).float()
            lo = (m - hi).to(torch.float16).float()
            m = torch.cat([hi, lo], dim=1)
        blocks.append(_pack_rows_sw128(m).reshape(-1))
    img = torch.cat(blocks).contiguous()
    params = torch.cat([v.float().reshape(-1) for v in (b_out, b1, b2, g1, be1, g2, be2)]).contiguous()
    assert params.numel() == 768
    return img, params


def pack_stem_tc(w, scale):
    """Operand image of i2r_stem_conv3x3s2_tc: w fp32 [K = Cin*9, 64] (k = (c*3+ky)*3+kx), BatchNorm scale folded in
    (fp32 product) and split into an fp16 pair.  Rows = output channels, 64 K slots of 128 bytes each:
    block 0 = [W_hi (32 slots, zero past K) | W_hi], block 1 = [W_lo | 0]; SWIZZLE_128B."""
    k, cout = w.shape
    assert cout == 64 and k <= 32
    m = (w.float() * scale.float().reshape(1, cout)).t().contiguous()      # [64, K]
    hi = m.to(torch.float16).float()
    lo = (m - hi).to(torch.float16).float()
    b1 = torch.zeros(cout, 64)
    b1[:, :k] = hi
    b1[:, 32:32 + k] = hi
    b2 = torch.zeros(cout, 64)
    b2[:, :k] = lo
    return torch.cat([_pack_rows_sw128(b1).reshape(-1), _pack_rows_sw128(b2).reshape(-1)]).contiguous()
