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
