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
        pos5 = mask_embedding_conv(sd, "position_embedding", pad_persons(pos_mask, length), list(m.TRANS_SIZE))
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
    if m.SINGLEFORMER != "transpose_h":
        raise NotImplementedError("oracle first stage %r" % m.SINGLEFORMER)
    feat, single = transpose_h_first_stage(sd, cfg, x, "singleformer.")
    if taps is not None:
        taps["feat"] = feat
    t = feat
    for _ in range(int(math.log(feat.shape[-1] // m.TRANS_SIZE[-1], 2))):      # interformer.py:260-264
        t = F.max_pool2d(t, 3, 2, 1)
    bs, n_max = len(length), max(length)
    _, c, h, w = t.shape
    pos = None
    if m.USE_MULTI_POS:
        if m.MULTI_POS_EMBEDDING != "conv":
            raise NotImplementedError("oracle MULTI_POS_EMBEDDING %r" % m.MULTI_POS_EMBEDDING)
        pos5 = mask_embedding_conv(sd, "multi_position_embedding", pad_persons(pos_mask, length), list(m.TRANS_SIZE))
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
    else:
        raise NotImplementedError("oracle UPSAMPLE_TYPE %r" % m.UPSAMPLE_TYPE)
    y = feat + y
    multi = F.conv2d(y, sd["final_layer.weight"], sd["final_layer.bias"])
    if m.INTER_SUPERVISION and not m.SINGLEFORMER_FIX:
        return {"single": single, "multi": multi}
    return multi


def forward(sd, cfg, x, pos_mask, length, taps=None):
    """Dispatch on cfg.MODEL.NAME like tools/test.py:87 does."""
    if cfg.MODEL.NAME == "interformer_pureMulti":
        return vanilla_forward(sd, cfg, x, pos_mask, length, taps)
    return two_stage_forward(sd, cfg, x, pos_mask, length, taps)
