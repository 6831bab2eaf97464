"""CPU oracle (TEST INFRASTRUCTURE): a plain torch-fp32 functional restatement of the I2R-Net forward
path, driven directly by a reference-format state_dict.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; it is the checker, never the product path.

Pinning: tests/golden/*.npz hold outputs of the REAL reference (imported from /root/reference through
oracle/ref_shims by tests/golden/make_golden.py) on synthetic weights/inputs; tests/test_oracle.py
checks this restatement against them (<= 2e-5 max-abs), so the oracle is pinned by reference outputs,
not by itself.  The reference has no golden vectors of its own.

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, BN_EPS)


def _conv(sd, p, x, stride=1):
    w = sd[p + ".weight"]
    return F.conv2d(x, w, sd.get(p + ".bias"), stride, w.shape[2] // 2)


def basic_block(sd, p, x):
    """lib/models/interformer_pureMulti.py:37-66"""
    out = F.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x)))
    out = _bn(sd, p + ".bn2", _conv(sd, p + ".conv2", out))
    res = x
    if (p + ".downsample.0.weight") in sd:
        res = _bn(sd, p + ".downsample.1", _conv(sd, p + ".downsample.0", x))
    return F.relu(out + res)


def bottleneck(sd, p, x):
    """lib/models/interformer_pureMulti.py:69-107"""
    out = F.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x)))
    out = F.relu(_bn(sd, p + ".bn2", _conv(sd, p + ".conv2", out)))
    out = _bn(sd, p + ".bn3", _conv(sd, p + ".conv3", out))
    res = x
    if (p + ".downsample.0.weight") in sd:
        res = _bn(sd, p + ".downsample.1", _conv(sd, p + ".downsample.0", x))
    return F.relu(out + res)


def _count(sd, prefix):
    """Number of consecutive integer children `prefix.0`, `prefix.1`, ... present in sd."""
    n = 0
    while any(k.startswith("%s.%d." % (prefix, n)) for k in sd):
        n += 1
    return n


def hr_module(sd, p, xs):
    """HighResolutionModule.forward + fuse layers (lib/models/interformer_pureMulti.py:332-410)."""
    nb = len(xs)
    xs = list(xs)
    for b in range(nb):
        for u in range(_count(sd, "%s.branches.%d" % (p, b))):
            xs[b] = basic_block(sd, "%s.branches.%d.%d" % (p, b, u), xs[b])
    if nb == 1:
        return xs
    outs = []
    for i in range(nb):
        y = None
        for j in range(nb):
            if j == i:
                t = xs[j]
            elif j > i:
                key = "%s.fuse_layers.%d.%d" % (p, i, j)
                t = _bn(sd, key + ".1", _conv(sd, key + ".0", xs[j]))
                t = F.interpolate(t, scale_factor=2 ** (j - i), mode="nearest")
            else:
                t = xs[j]
                for s in range(i - j):
                    key = "%s.fuse_layers.%d.%d.%d" % (p, i, j, s)
                    t = _bn(sd, key + ".1", _conv(sd, key + ".0", t, stride=2))
                    if s < i - j - 1:
                        t = F.relu(t)
            y = t if y is None else y + t
        outs.append(F.relu(y))
    return outs


def hrnet_w48s_backbone(sd, x, prefix=""):
    """Stem, layer1, transition1, stage2, transition2, stage3 (interformer_pureMulti.py:675-699)."""
    P = prefix
    x = F.relu(_bn(sd, P + "bn1", _conv(sd, P + "conv1", x, 2)))
    x = F.relu(_bn(sd, P + "bn2", _conv(sd, P + "conv2", x, 2)))
    for u in range(_count(sd, P + "layer1")):
        x = bottleneck(sd, "%slayer1.%d" % (P, u), x)

    def transition(name, feats, nbranch):
        out = []
        for i in range(nbranch):
            base = "%s%s.%d" % (P, name, i)
            if (base + ".0.weight") in sd:                       # same-resolution 3x3 + BN + ReLU
                out.append(F.relu(_bn(sd, base + ".1", _conv(sd, base + ".0", feats[min(i, len(feats) - 1)]))))
            elif (base + ".0.0.weight") in sd:                   # new branch: chain of stride-2 3x3
                t = feats[-1]
                for s in range(_count(sd, base)):
                    t = F.relu(_bn(sd, "%s.%d.1" % (base, s), _conv(sd, "%s.%d.0" % (base, s), t, 2)))
                out.append(t)
            else:
                out.append(feats[i])
        return out

    nb2 = _count(sd, P + "stage2.0.branches")
    xs = transition("transition1", [x], nb2)
    for m in range(_count(sd, P + "stage2")):
        xs = hr_module(sd, "%sstage2.%d" % (P, m), xs)
    nb3 = _count(sd, P + "stage3.0.branches")
    xs = transition("transition2", xs, nb3)
    for m in range(_count(sd, P + "stage3")):
        xs = hr_module(sd, "%sstage3.%d" % (P, m), xs)
    return xs


def pad_persons(t, length):
    """padding_tensor (interformer_pureMulti.py:721-742): [S,...] -> [bs, N, ...], zero rows for missing persons."""
    n_max = max(length)
    rows = []
    for item in t.split(list(length), dim=0):
        if item.shape[0] < n_max:
            item = torch.cat([item, item.new_zeros((n_max - item.shape[0],) + tuple(item.shape[1:]))], dim=0)
        rows.append(item)
    return torch.stack(rows, dim=0)


def person_mask(length, hw):
    """get_mask (interformer_pureMulti.py:706-719): bool [bs, N, h, w], True on padded persons."""
    n_max = max(length)
    m = torch.zeros((len(length), n_max) + tuple(hw), dtype=torch.bool)
    for b, n in enumerate(length):
        m[b, n:] = True
    return m


def mask_embedding_conv(sd, p, pos_mask5, trans_size):
    """PositionEmbeddingImage.forward, mode 'conv' (lib/models/position_embedding.py:65-116)."""
    bs, n, c, h, w = pos_mask5.shape
    x = pos_mask5.reshape(bs * n, c, h, w)
    x = F.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x, 2)))
    x = F.relu(_bn(sd, p + ".bn2", _conv(sd, p + ".conv2", x, 2)))
    for _ in range(int(math.log(x.shape[-1] // trans_size[-1], 2))):
        x = F.max_pool2d(x, 3, 2, 1)
    return x.reshape(bs, n, x.shape[-3], x.shape[-2], x.shape[-1])


def mask_embedding_res(sd, p, pos_mask5, trans_size):
    """PositionEmbeddingImage.forward, mode 'res' (lib/models/position_embedding.py:14-18, :90-108): conv_pre 3x3 (1 -> 3),
    the first five children of torchvision resnet18 (conv1 7x7 s2 p3, bn1, relu, maxpool 3/2/1, layer1 = two BasicBlocks),
    conv_end 3x3 (64 -> d), then max-pools down to TRANS_SIZE."""
    bs, n, c, h, w = pos_mask5.shape
    x = pos_mask5.reshape(bs * n, c, h, w)
    x = _conv(sd, p + ".conv_pre", x)
    x = F.relu(_bn(sd, p + ".res.1", F.conv2d(x, sd[p + ".res.0.weight"], None, 2, 3)))
    x = F.max_pool2d(x, 3, 2, 1)
    for b in (0, 1):
        x = basic_block(sd, "%s.res.4.%d" % (p, b), x)
    x = _conv(sd, p + ".conv_end", x)
    for _ in range(int(math.log(x.shape[-1] // trans_size[-1], 2))):
        x = F.max_pool2d(x, 3, 2, 1)
    return x.reshape(bs, n, x.shape[-3], x.shape[-2], x.shape[-1])


def mask_embedding(sd, p, pos_mask5, trans_size, mode):
    if mode == "conv":
        return mask_embedding_conv(sd, p, pos_mask5, trans_size)
    if mode == "res":
        return mask_embedding_res(sd, p, pos_mask5, trans_size)
    raise NotImplementedError("MULTI_POS_EMBEDDING=%r" % mode)


def mha_single_head(sd, p, q_in, k_in, v_in, key_padding_mask):
    """nn.MultiheadAttention(nhead=1) forward as torch implements it (torch/nn/functional.py
    multi_head_attention_forward): packed in_proj, q * head_dim**-0.5, -inf on padded keys, softmax,
    PV, out_proj.  Inputs [L, B, E]; mask bool [B, L]."""
    w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    e = q_in.shape[-1]
    q = F.linear(q_in, w[:e], b[:e]) * (float(e) ** -0.5)
    k = F.linear(k_in, w[e:2 * e], b[e:2 * e])
    v = F.linear(v_in, w[2 * e:], b[2 * e:])
    q, k, v = q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1)       # [B, L, E]
    scores = torch.bmm(q, k.transpose(1, 2))
    if key_padding_mask is not None:
        scores = scores.masked_fill(key_padding_mask[:, None, :], float("-inf"))
    attn = torch.softmax(scores, dim=-1)
    out = torch.bmm(attn, v).transpose(0, 1)                                  # [L, B, E]
    return F.linear(out, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def encoder_post_norm(sd, p, src, pos, key_padding_mask, num_layers):
    """TransformerEncoder of forward_post layers (interformer_pureMulti.py:125-148, :182-213)."""
    for i in range(num_layers):
        lp = "%s.layers.%d" % (p, i)
        qk = src if pos is None else src + pos
        a = mha_single_head(sd, lp + ".self_attn", qk, qk, src, key_padding_mask)
        src = F.layer_norm(src + a, (src.shape[-1],), sd[lp + ".norm1.weight"], sd[lp + ".norm1.bias"], 1e-5)
        f = F.linear(F.relu(F.linear(src, sd[lp + ".linear1.weight"], sd[lp + ".linear1.bias"])),
                     sd[lp + ".linear2.weight"], sd[lp + ".linear2.bias"])
        src = F.layer_norm(src + f, (src.shape[-1],), sd[lp + ".norm2.weight"], sd[lp + ".norm2.bias"], 1e-5)
    return src


def unpad_persons(t, length):
    """get_valid_output (lib/utils/utils.py:24-37)."""
    n_max = max(length)
    g = t.reshape((t.shape[0] // n_max, n_max) + tuple(t.shape[1:]))
    return torch.cat([g[i, :n] for i, n in enumerate(length)], dim=0)


def vanilla_forward(sd, cfg, x, pos_mask, length, taps=None):
    """interformer_pureMulti.TransPoseH.forward (lib/models/interformer_pureMulti.py:752-778).

    `taps` (optional dict) receives intermediate tensors for debugging the CUDA path.
    """
    sd = {k: v.float() for k, v in sd.items() if v.dtype.is_floating_point}
    m = cfg.MODEL
    length = list(length)
    feats = hrnet_w48s_backbone(sd, x)
    x = F.conv2d(feats[-1], sd["reduce.weight"])                                # [S, d, 16, 12]
    if taps is not None:
        taps["branches"] = feats
        taps["reduce"] = x
    bs, n_max = len(length), max(length)
    _, c, h, w = x.shape
    pos = None
    if m.USE_MULTI_POS:
        pos5 = mask_embedding(sd, "position_embedding", pad_persons(pos_mask, length), list(m.TRANS_SIZE),
                              m.MULTI_POS_EMBEDDING)
        if taps is not None:
            taps["pos"] = unpad_persons(pos5.reshape((bs * n_max,) + tuple(pos5.shape[2:])), length)
        pos = pos5.permute(0, 2, 1, 3, 4).flatten(2).permute(2, 0, 1)          # [N*h*w, bs, c]
    mask = person_mask(length, (h, w)).flatten(1)
    src = pad_persons(x, length).permute(0, 2, 1, 3, 4).flatten(2).permute(2, 0, 1)
    y = encoder_post_norm(sd, "global_encoder", src, pos, mask, m.ENCODER_LAYERS)
    y = y.permute(1, 2, 0).contiguous().view(bs, c, n_max, h, w)
    y = y.permute(0, 2, 1, 3, 4).contiguous().view(bs * n_max, c, h, w)
    if taps is not None:
        taps["encoded"] = unpad_persons(y, length)
    for _ in range(2):                                                          # same deconv stack twice (:774-775)
        for i in range(m.EXTRA.NUM_DECONV_LAYERS):
            y = F.conv_transpose2d(y, sd["deconv_layers.%d.weight" % (3 * i)],
                                   sd.get("deconv_layers.%d.bias" % (3 * i)), 2, 1, 0)
            y = F.relu(_bn(sd, "deconv_layers.%d" % (3 * i + 1), y))
    y = F.conv2d(y, sd["final_layer.weight"], sd["final_layer.bias"])
    return unpad_persons(y, length)


def transpose_h_first_stage(sd, cfg, x, prefix=""):
    """transpose_h.TransPoseH.forward (lib/models/transpose_h.py:623-655): backbone -> reduce on branch
    HRNET_RES_LAYER -> post-norm encoder over the h*w tokens of each crop, pos = the [h*w,1,d] parameter ->
    (feature map, final_layer heatmaps).  `sd` holds fp32 tensors; keys are looked up under `prefix`."""
    m = cfg.MODEL
    feats = hrnet_w48s_backbone(sd, x, prefix)
    f = F.conv2d(feats[int(m.HRNET_RES_LAYER)], sd[prefix + "reduce.weight"])
    bs, c, h, w = f.shape
    src = f.flatten(2).permute(2, 0, 1)                                         # [h*w, S, c]
    pos = sd.get(prefix + "pos_embedding")
    y = encoder_post_norm(sd, prefix + "global_encoder", src, pos, None, m.ENCODER_LAYERS)
    feat = y.permute(1, 2, 0).contiguous().view(bs, c, h, w)
    return feat, F.conv2d(feat, sd[prefix + "final_layer.weight"], sd[prefix + "final_layer.bias"])


# ------------------------------------------------------------------------------------------------ HRFormer-B first stage
def _hrt_window_attention(sd, p, x, heads, ws=7):
    """InterlacedPoolAttention.forward + MHA_.forward (lib/models/hrformer.py:1164-1180, :627-935) on x [B, H, W, C]
    (already LayerNorm-ed): center zero-pad H, W to multiples of 7 (:949-956), 7x7 windows as sequences
    (:976-986), q/k/v = three biased Linear(C, C), q * head_dim**-0.5, softmax(q k^T) WITHOUT the relative position
    bias (it is computed but its addition is commented out, :866-888) and without any mask -- padded tokens are zero
    vectors, so their keys/values are the projection biases and they DO take part --, P v, out_proj, de-pad."""
    b, h, w, c = x.shape
    ph, pw = (-h) % ws, (-w) % ws
    xp = F.pad(x, (0, 0, pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
    hp, wp = h + ph, w + pw
    qh, qw = hp // ws, wp // ws
    # "n (qh ph) (qw pw) c -> (ph pw) (n qh qw) c"
    t = xp.view(b, qh, ws, qw, ws, c).permute(2, 4, 0, 1, 3, 5).reshape(ws * ws, b * qh * qw, c)
    hd = c // heads
    q = F.linear(t, sd[p + ".q_proj.weight"], sd[p + ".q_proj.bias"]) * (float(hd) ** -0.5)
    k = F.linear(t, sd[p + ".k_proj.weight"], sd[p + ".k_proj.bias"])
    v = F.linear(t, sd[p + ".v_proj.weight"], sd[p + ".v_proj.bias"])
    n, bw, _ = q.shape
    q = q.contiguous().view(n, bw * heads, hd).transpose(0, 1)
    k = k.contiguous().view(n, bw * heads, hd).transpose(0, 1)
    v = v.contiguous().view(n, bw * heads, hd).transpose(0, 1)
    a = torch.softmax(torch.bmm(q, k.transpose(1, 2)), dim=-1)
    o = torch.bmm(a, v).transpose(0, 1).contiguous().view(n, bw, c)
    o = F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])
    # "(ph pw) (n qh qw) c -> n (qh ph) (qw pw) c"
    o = o.view(ws, ws, b, qh, qw, c).permute(2, 3, 0, 4, 1, 5).reshape(b, hp, wp, c)
    return o[:, ph // 2: ph // 2 + h, pw // 2: pw // 2 + w, :]


def _hrt_block(sd, p, x, heads):
    """GeneralTransformerBlock.forward (lib/models/hrformer.py:1230-1240) with MlpDWBN (:1094-1119): pre-norm
    (LayerNorm eps 1e-6), x + attn(norm1 x), x + mlp(norm2 x); mlp = 1x1 conv + BN + GELU, depthwise 3x3 + BN + GELU,
    1x1 conv + BN + GELU (drop_path / dropout are identity in eval)."""
    b, c, h, w = x.shape
    t = x.flatten(2).permute(0, 2, 1)                                                 # [B, HW, C]
    n1 = F.layer_norm(t, (c,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-6)
    t = t + _hrt_window_attention(sd, p + ".attn.attn", n1.view(b, h, w, c), heads).reshape(b, h * w, c)
    n2 = F.layer_norm(t, (c,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-6)
    m = n2.permute(0, 2, 1).reshape(b, c, h, w)
    m = F.gelu(_bn(sd, p + ".mlp.norm1", F.conv2d(m, sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"])))
    m = F.gelu(_bn(sd, p + ".mlp.norm2", F.conv2d(m, sd[p + ".mlp.dw3x3.weight"], sd[p + ".mlp.dw3x3.bias"], 1, 1, 1,
                                                  m.shape[1])))
    m = F.gelu(_bn(sd, p + ".mlp.norm3", F.conv2d(m, sd[p + ".mlp.fc2.weight"], sd[p + ".mlp.fc2.bias"])))
    t = t + m.flatten(2).permute(0, 2, 1)
    return t.permute(0, 2, 1).reshape(b, c, h, w)


def _hrt_module(sd, p, xs, heads, multiscale_output=True):
    """HighResolutionTransformerModule.forward (lib/models/hrformer.py:1708-1732) with _make_fuse_layers (:1616-1706):
    branches of transformer blocks, then for every output branch i: sum_j f_ij(x_j), ReLU; f_ij = 1x1 conv + BN +
    bilinear upsample x2^(j-i) (align_corners=False) for j > i, identity for j = i, a chain of [depthwise 3x3 s2 + BN +
    1x1 conv + BN (+ ReLU except the last)] for j < i."""
    nb = len(xs)
    ys = []
    for i in range(nb):
        y = xs[i]
        for blk in range(_count(sd, "%s.branches.%d" % (p, i))):
            y = _hrt_block(sd, "%s.branches.%d.%d" % (p, i, blk), y, heads[i])
        ys.append(y)
    if nb == 1:
        return ys
    out = []
    for i in range(nb if multiscale_output else 1):
        acc = None
        for j in range(nb):
            fp = "%s.fuse_layers.%d.%d" % (p, i, j)
            if j == i:
                t = ys[j]
            elif j > i:
                t = _bn(sd, fp + ".1", F.conv2d(ys[j], sd[fp + ".0.weight"]))
                t = F.interpolate(t, scale_factor=2 ** (j - i), mode="bilinear", align_corners=False)
                t = F.interpolate(t, size=ys[i].shape[2:], mode="bilinear", align_corners=False)     # `resize` (:1723)
            else:
                t = ys[j]
                for k in range(i - j):
                    kp = "%s.%d" % (fp, k)
                    t = _bn(sd, kp + ".1", F.conv2d(t, sd[kp + ".0.weight"], None, 2, 1, 1, t.shape[1]))
                    t = _bn(sd, kp + ".3", F.conv2d(t, sd[kp + ".2.weight"]))
                    if k != i - j - 1:
                        t = F.relu(t)
            acc = t if acc is None else acc + t
        out.append(F.relu(acc))
    return out


HRT_STAGES = (                # lib/models/hrformer.py:2489-2525 (hard-coded in get_pose_net)
    ("stage2", 1, (2, 4)),
    ("stage3", 4, (2, 4, 8)),
    ("stage4", 2, (2, 4, 8, 16)),
)


def hrformer_first_stage(sd, cfg, x, prefix=""):
    """hrformer.HRFormer.forward (lib/models/hrformer.py:2477-2480) = HRT.forward (:2057-2092) + TopDownSimpleHead
    with zero deconvs (:2343-2348): stem (two 3x3 s2 conv + BN + ReLU), two Bottlenecks, three transformer stages
    (1 / 4 / 2 modules; the last module of stage 4 only produces branch 0), 1x1 head on branch 0.  Returns
    (branch 0 feature [S, 78, H/4, W/4], heatmaps)."""
    b = prefix + "backbone."
    y = F.relu(_bn(sd, b + "bn1", _conv(sd, b + "conv1", x, 2)))
    y = F.relu(_bn(sd, b + "bn2", _conv(sd, b + "conv2", y, 2)))
    for i in range(_count(sd, b + "layer1")):
        y = bottleneck(sd, "%slayer1.%d" % (b, i), y)

    def transition(name, idx, t):
        tp = "%s%s.%d" % (b, name, idx)
        if (tp + ".0.weight") in sd and sd[tp + ".0.weight"].dim() == 4:     # 3x3 s1 conv + BN + ReLU (:1874-1890)
            return F.relu(_bn(sd, tp + ".1", _conv(sd, tp + ".0", t)))
        for j in range(_count(sd, tp)):                                         # 3x3 s2 conv + BN + ReLU chain
            t = F.relu(_bn(sd, "%s.%d.1" % (tp, j), _conv(sd, "%s.%d.0" % (tp, j), t, 2)))
        return t
    xs = [transition("transition1", 0, y), transition("transition1", 1, y)]
    for si, (stage, nmod, heads) in enumerate(HRT_STAGES):
        if si > 0:
            xs = xs + [transition("transition%d" % (si + 1), len(xs), xs[-1])]
        for mi in range(nmod):
            last = stage == "stage4" and mi == nmod - 1
            xs = _hrt_module(sd, "%s%s.%d" % (b, stage, mi), xs, heads, multiscale_output=not last)
    feat = xs[0]
    hp = prefix + "keypoint_head.final_layer"
    return feat, F.conv2d(feat, sd[hp + ".weight"], sd[hp + ".bias"])


def _deconv_block(sd, key, y, num_layers):
    for i in range(num_layers):
        y = F.conv_transpose2d(y, sd["%s.%d.weight" % (key, 3 * i)], sd.get("%s.%d.bias" % (key, 3 * i)), 2, 1, 0)
        y = F.relu(_bn(sd, "%s.%d" % (key, 3 * i + 1), y))
    return y


def two_stage_forward(sd, cfg, x, pos_mask, length, taps=None):
    """InterFormer.forward of lib/models/interformer.py:282-323 and lib/models/interformer_2stage.py:383-423
    (same arithmetic; cfg.MODEL.NAME selects the parameter names of the upsample path) with a TransPose-H first
    stage.  Returns {'single', 'multi'} or the 'multi' tensor, as the reference does."""
    sd = {k: v.float() for k, v in sd.items() if v.dtype.is_floating_point}
    m = cfg.MODEL
    length = list(length)
    if m.SINGLEFORMER == "transpose_h":
        feat, single = transpose_h_first_stage(sd, cfg, x, "singleformer.")
    elif m.SINGLEFORMER == "hrformer":
        feat, single = hrformer_first_stage(sd, cfg, x, "singleformer.")
    elif not m.SINGLEFORMER:
        # stand-alone HRNet backbone (interformer.py:143, :291-292; lib/models/hrnet.py:413-446): token map directly
        feat, single = None, None
        t = F.conv2d(hrnet_w48s_backbone(sd, x, "backbone.body.")[-1], sd["backbone.body.reduce.weight"])
    else:
        raise NotImplementedError("oracle first stage %r" % m.SINGLEFORMER)
    if taps is not None:
        taps["feat"] = feat
    if feat is not None:
        t = feat
        for _ in range(int(math.log(feat.shape[-1] // m.TRANS_SIZE[-1], 2))):      # interformer.py:260-264
            t = F.max_pool2d(t, 3, 2, 1)
    bs, n_max = len(length), max(length)
    _, c, h, w = t.shape
    pos = None
    if m.USE_MULTI_POS:
        pos5 = mask_embedding(sd, "multi_position_embedding", pad_persons(pos_mask, length), list(m.TRANS_SIZE),
                              m.MULTI_POS_EMBEDDING)
        pos = pos5.permute(0, 2, 1, 3, 4).flatten(2).permute(2, 0, 1)
    mask = person_mask(length, (h, w)).flatten(1)
    src = pad_persons(t, length).permute(0, 2, 1, 3, 4).flatten(2).permute(2, 0, 1)
    y = encoder_post_norm(sd, "multi_global_encoder", src, pos, mask, m.ENCODER_MULTI_LAYERS)
    y = y.permute(1, 2, 0).contiguous().view(bs, c, n_max, h, w)
    y = unpad_persons(y.permute(0, 2, 1, 3, 4).reshape(bs * n_max, c, h, w), length)
    if taps is not None:
        taps["encoded"] = y
    nl = m.EXTRA.NUM_DECONV_LAYERS
    steps = int(math.log(m.HEATMAP_SIZE[0] // m.TRANS_SIZE[1], 2))
    if m.UPSAMPLE_TYPE == "multiplex":
        # interformer.py applies the stack twice (:310-312); interformer_2stage log2(ratio) times (:375-377)
        for _ in range(steps if m.NAME == "interformer_2stage" else 2):
            y = _deconv_block(sd, "deconv_layers", y, nl)
    elif m.UPSAMPLE_TYPE == "deconv":
        for i in range(steps):
            key = "deconv_layers%d" % (i + 1) if m.NAME == "interformer_2stage" else "upsample_layer.deconv_layers.%d" % i
            y = _deconv_block(sd, key, y, nl)
    elif m.UPSAMPLE_TYPE == "upconv":      # interformer.py:25-64
        u = "upsample_layer."
        y = _bn(sd, u + "fuse_layers.1", _conv(sd, u + "fuse_layers.0", y))
        y = F.interpolate(y, scale_factor=m.HEATMAP_SIZE[0] // m.TRANS_SIZE[1], mode="nearest")
        y = F.relu(_bn(sd, u + "double_conv.1", _conv(sd, u + "double_conv.0", y)))
        y = F.relu(_bn(sd, u + "double_conv.4", _conv(sd, u + "double_conv.3", y)))
    else:
        raise NotImplementedError("oracle UPSAMPLE_TYPE %r" % m.UPSAMPLE_TYPE)
    if feat is not None:
        y = feat + y
    multi = F.conv2d(y, sd["final_layer.weight"], sd["final_layer.bias"])
    if m.INTER_SUPERVISION and m.SINGLEFORMER and not m.SINGLEFORMER_FIX:
        return {"single": single, "multi": multi}
    return multi


def forward(sd, cfg, x, pos_mask, length, taps=None):
    """Dispatch on cfg.MODEL.NAME like tools/test.py:87 does."""
    if cfg.MODEL.NAME == "interformer_pureMulti":
        return vanilla_forward(sd, cfg, x, pos_mask, length, taps)
    return two_stage_forward(sd, cfg, x, pos_mask, length, taps)
