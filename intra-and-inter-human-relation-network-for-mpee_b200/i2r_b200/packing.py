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


def pick_kc(cin):
    if cin % 64 == 0:
        return 64
    if cin % 48 == 0:
        return 48
    raise ValueError("Cin=%d is not a multiple of 48 or 64 (channel padding not implemented)" % cin)


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


def pack_taps(mats, kc):
    """mats: list over taps of fp32 [Cout, Cin] matrices -> fp16 [ntaps, Cin/kc, Npad, 64].

    Per (tap, K-chunk) block: Npad rows (output channels) of 128 B = 64 fp16 K-slots, of which the first
    `kc` hold input channels and the rest are zero; the eight 16-byte chunks of row n are stored at chunk
    position (c XOR (n & 7)) -- the SWIZZLE_128B K-major image the kernels bulk-copy into shared memory.
    """
    cout, cin = mats[0].shape
    npad = ceil_to(cout, 16)
    nch, kg = cin // kc, kc // 8
    out = torch.zeros(len(mats), nch, npad, 8, 8, dtype=torch.float16)
    rows = torch.arange(npad)
    for t, m in enumerate(mats):
        mm = m.float().reshape(cout, nch, kg, 8).permute(1, 0, 2, 3).to(torch.float16)   # [nch, cout, kg, 8]
        for c in range(kg):
            pos = (c ^ (rows[:cout] & 7))
            out[t, :, rows[:cout], pos, :] = mm[:, :, c, :]
    return out.reshape(len(mats), nch, npad, 64).contiguous()


def unpack_taps(packed, kc):
    """Inverse of pack_taps: fp16 [ntaps, nch, Npad, 64] -> fp32 [ntaps, Npad, Cin] (tests / emulator)."""
    ntaps, nch, npad, _ = packed.shape
    kg = kc // 8
    p5 = packed.reshape(ntaps, nch, npad, 8, 8).float()
    rows = torch.arange(npad)
    out = torch.zeros(ntaps, nch, npad, kg, 8)
    for c in range(kg):
        pos = (c ^ (rows & 7))
        out[:, :, rows, c, :] = p5[:, :, rows, pos, :]
    return out.permute(0, 2, 1, 3, 4).reshape(ntaps, npad, nch * kc)


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
