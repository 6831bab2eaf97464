"""HRNet-W48-S backbone (stem, layer1, two multi-resolution stages) for the B200 path.

Two halves:
  * `attach_backbone_params(module, extra)` registers parameter holders on an nn.Module under the
    reference's state_dict names (reference: lib/models/interformer_pureMulti.py:427-455 and the
    identical blocks in transpose_h.py / hrnet.py) so reference checkpoints load with strict=True.
    The holders never run a torch forward.
  * `BackboneProgram` folds BatchNorm, packs weights once, and executes the backbone as grouped
    implicit-GEMM launches (all resolution branches of a module share one grid).
"""
import torch
import torch.nn as nn

from .ops import ConvLayer
from .packing import conv_taps, fold_bn


# --------------------------------------------------------------------------- parameter holders
def _bn(c):
    return nn.BatchNorm2d(c, momentum=0.1)


def _conv(cin, cout, k, stride=1):
    return nn.Conv2d(cin, cout, k, stride, k // 2, bias=False)


class ResidualUnitParams(nn.Module):
    """conv1/bn1/conv2/bn2[/conv3/bn3][/downsample] holder (BasicBlock: expansion 1, Bottleneck: 4)."""

    def __init__(self, kind, cin, planes, with_downsample):
        super().__init__()
        self.kind = kind
        if kind == "BASIC":
            self.conv1, self.bn1 = _conv(cin, planes, 3), _bn(planes)
            self.conv2, self.bn2 = _conv(planes, planes, 3), _bn(planes)
            cout = planes
        elif kind == "BOTTLENECK":
            self.conv1, self.bn1 = _conv(cin, planes, 1), _bn(planes)
            self.conv2, self.bn2 = _conv(planes, planes, 3), _bn(planes)
            self.conv3, self.bn3 = _conv(planes, planes * 4, 1), _bn(planes * 4)
            cout = planes * 4
        else:
            raise ValueError("unknown block type %r" % kind)
        self.downsample = nn.Sequential(_conv(cin, cout, 1), _bn(cout)) if with_downsample else None
        self.out_channels = cout


EXPANSION = {"BASIC": 1, "BOTTLENECK": 4}


def _unit_chain(kind, cin, planes, count):
    units, c = [], cin
    for i in range(count):
        u = ResidualUnitParams(kind, c, planes, with_downsample=(i == 0 and c != planes * EXPANSION[kind]))
        units.append(u)
        c = u.out_channels
    return nn.Sequential(*units), c


class HRModuleParams(nn.Module):
    """`branches` + `fuse_layers` holder of one multi-resolution module."""

    def __init__(self, kind, num_blocks, in_channels, channels, multi_scale_output=True):
        super().__init__()
        nb = len(channels)
        if not (nb == len(num_blocks) == len(in_channels)):
            raise ValueError("NUM_BRANCHES(%d) <> NUM_BLOCKS(%d) / NUM_CHANNELS / NUM_INCHANNELS" % (
                nb, len(num_blocks)))
        branches, outs = [], []
        for b in range(nb):
            seq, c = _unit_chain(kind, in_channels[b], channels[b], num_blocks[b])
            branches.append(seq)
            outs.append(c)
        self.branches = nn.ModuleList(branches)
        self.out_channels = outs
        rows = []
        for i in range(nb if multi_scale_output else 1):
            row = []
            for j in range(nb):
                if j > i:    # lower resolution -> 1x1 conv, BN, nearest upsample
                    row.append(nn.Sequential(_conv(outs[j], outs[i], 1), _bn(outs[i]),
                                             nn.Upsample(scale_factor=2 ** (j - i), mode="nearest")))
                elif j == i:
                    row.append(None)
                else:        # higher resolution -> chain of stride-2 3x3 convs
                    chain = []
                    for s in range(i - j):
                        last = s == i - j - 1
                        co = outs[i] if last else outs[j]
                        mods = [_conv(outs[j], co, 3, 2), _bn(co)] + ([] if last else [nn.ReLU(True)])
                        chain.append(nn.Sequential(*mods))
                    row.append(nn.Sequential(*chain))
            rows.append(nn.ModuleList(row))
        self.fuse_layers = nn.ModuleList(rows) if nb > 1 else None


def _transition(pre, cur):
    layers = []
    for i, c in enumerate(cur):
        if i < len(pre):
            layers.append(None if c == pre[i] else nn.Sequential(_conv(pre[i], c, 3), _bn(c), nn.ReLU(True)))
        else:
            chain = []
            for s in range(i + 1 - len(pre)):
                co = c if s == i - len(pre) else pre[-1]
                chain.append(nn.Sequential(_conv(pre[-1], co, 3, 2), _bn(co), nn.ReLU(True)))
            layers.append(nn.Sequential(*chain))
    return nn.ModuleList(layers)


def _stage(cfg_stage, in_channels):
    kind = cfg_stage["BLOCK"]
    mods = []
    chans = list(in_channels)
    for _ in range(cfg_stage["NUM_MODULES"]):
        m = HRModuleParams(kind, list(cfg_stage["NUM_BLOCKS"]), chans, list(cfg_stage["NUM_CHANNELS"]))
        chans = m.out_channels
        mods.append(m)
    return nn.Sequential(*mods), chans


def attach_backbone_params(module, extra):
    """Register conv1..stage3 on `module`; returns the channel list of the last stage's branches."""
    module.conv1, module.bn1 = _conv(3, 64, 3, 2), _bn(64)
    module.conv2, module.bn2 = _conv(64, 64, 3, 2), _bn(64)
    module.layer1, c1 = _unit_chain("BOTTLENECK", 64, 64, 4)
    s2, s3 = extra["STAGE2"], extra["STAGE3"]
    ch2 = [c * EXPANSION[s2["BLOCK"]] for c in s2["NUM_CHANNELS"]]
    module.transition1 = _transition([c1], ch2)
    module.stage2, pre = _stage(s2, ch2)
    ch3 = [c * EXPANSION[s3["BLOCK"]] for c in s3["NUM_CHANNELS"]]
    module.transition2 = _transition(pre, ch3)
    module.stage3, pre = _stage(s3, ch3)
    return pre


# --------------------------------------------------------------------------- compiled program
def conv_bn_layer(sd, conv_key, bn_key, stride=1, relu=False, device="cuda"):
    w = sd[conv_key + ".weight"].float()
    mats, dys, dxs = conv_taps(w, pad=w.shape[2] // 2)
    if bn_key is not None:
        scale, bias = fold_bn(sd, bn_key, w.shape[0], conv_bias=sd.get(conv_key + ".bias"))
    else:
        scale = torch.ones(w.shape[0])
        b = sd.get(conv_key + ".bias")
        bias = b.float() if b is not None else torch.zeros(w.shape[0])
    return ConvLayer(mats, dys, dxs, scale, bias, stride=stride, relu=relu, device=device)


class _Unit:
    def __init__(self, sd, prefix, kind, device):
        self.kind = kind
        self.c1 = conv_bn_layer(sd, prefix + ".conv1", prefix + ".bn1", relu=True, device=device)
        if kind == "BASIC":
            self.c2 = conv_bn_layer(sd, prefix + ".conv2", prefix + ".bn2", relu=True, device=device)
        else:
            self.c2 = conv_bn_layer(sd, prefix + ".conv2", prefix + ".bn2", relu=True, device=device)
            self.c3 = conv_bn_layer(sd, prefix + ".conv3", prefix + ".bn3", relu=True, device=device)
        self.ds = None
        if (prefix + ".downsample.0.weight") in sd:
            self.ds = conv_bn_layer(sd, prefix + ".downsample.0", prefix + ".downsample.1", device=device)


class _HRModule:
    def __init__(self, sd, prefix, params, device):
        self.nb = len(params.branches)
        self.units = [[_Unit(sd, "%s.branches.%d.%d" % (prefix, b, u), params.branches[b][u].kind, device)
                       for u in range(len(params.branches[b]))] for b in range(self.nb)]
        self.fuse = {}
        for i, row in enumerate(params.fuse_layers):
            for j, f in enumerate(row):
                if f is None:
                    continue
                key = "%s.fuse_layers.%d.%d" % (prefix, i, j)
                if j > i:
                    self.fuse[(i, j)] = [conv_bn_layer(sd, key + ".0", key + ".1", device=device)]
                else:
                    self.fuse[(i, j)] = [
                        conv_bn_layer(sd, "%s.%d.0" % (key, s), "%s.%d.1" % (key, s), stride=2,
                                      relu=(s < i - j - 1), device=device) for s in range(i - j)]


class BackboneProgram:
    """Executes conv1 .. stage3 on a Runner; returns the list of branch feature maps (fp16 NHWC)."""

    def __init__(self, model, sd, device):
        w = sd["conv1.weight"].float()                       # [64,3,3,3] -> [27,64], k=(c*3+ky)*3+kx
        self.stem_w = w.permute(1, 2, 3, 0).reshape(-1, w.shape[0]).contiguous().to(device)
        sc, bi = fold_bn(sd, "bn1", w.shape[0])
        self.stem_scale, self.stem_bias = sc.to(device), bi.to(device)
        self.conv2 = conv_bn_layer(sd, "conv2", "bn2", stride=2, relu=True, device=device)
        self.layer1 = [_Unit(sd, "layer1.%d" % i, "BOTTLENECK", device) for i in range(len(model.layer1))]
        self.trans1 = self._transition(sd, "transition1", model.transition1, device)
        self.stage2 = [_HRModule(sd, "stage2.%d" % i, m, device) for i, m in enumerate(model.stage2)]
        self.trans2 = self._transition(sd, "transition2", model.transition2, device)
        self.stage3 = [_HRModule(sd, "stage3.%d" % i, m, device) for i, m in enumerate(model.stage3)]

    @staticmethod
    def _transition(sd, prefix, params, device):
        out = []
        for i, t in enumerate(params):
            if t is None:
                out.append(None)
            elif isinstance(t[0], nn.Conv2d):
                out.append([conv_bn_layer(sd, "%s.%d.0" % (prefix, i), "%s.%d.1" % (prefix, i), relu=True,
                                          device=device)])
            else:
                out.append([conv_bn_layer(sd, "%s.%d.%d.0" % (prefix, i, s), "%s.%d.%d.1" % (prefix, i, s),
                                          stride=2, relu=True, device=device) for s in range(len(t))])
        return out

    # ---- execution
    @staticmethod
    def _bottleneck(r, u, x):
        h = r.conv(u.c1, x)
        h = r.conv(u.c2, h)
        res = r.conv(u.ds, x) if u.ds is not None else x
        return r.conv(u.c3, h, add0=res)

    @staticmethod
    def _branches(r, mod, xs):
        depth = len(mod.units[0])
        for u in range(depth):
            hs = r.conv_group([(mod.units[b][u].c1, xs[b], {}) for b in range(mod.nb)])
            xs = r.conv_group([(mod.units[b][u].c2, hs[b], {"add0": xs[b]}) for b in range(mod.nb)])
        return xs

    @staticmethod
    def _fuse(r, mod, xs):
        """Multi-resolution exchange: y_i = relu(sum_j f_ij(x_j))  (reference :392-410).

        The lower-resolution output branches end in ONE implicit GEMM (the stride-2 3x3 from the branch above) whose
        epilogue adds the identity branch and the remaining terms through (up-sampling) addends.  The highest
        resolution branch has no convolution of its own left: a 1x1 convolution commutes with nearest upsampling, so
        its 1x1 terms run at THEIR resolution (4x / 16x fewer MMAs than at the output resolution) and one
        elementwise pass sums identity + upsampled terms and applies the ReLU.
        """
        f = mod.fuse
        if mod.nb == 2:
            t01, y1 = r.conv_group([
                (f[(0, 1)][0], xs[1], {}),                                  # 1x1 96->48 at 1/8 (commutes with upsampling)
                (f[(1, 0)][0], xs[0], dict(add0=xs[1], relu=True)),          # 3x3 s2 48->96 + identity
            ])
            return [r.upsum(xs[0], t01, 1), y1]
        if mod.nb != 3:
            raise NotImplementedError("fuse for %d resolution branches" % mod.nb)
        t01, t02, t12, u20, t21 = r.conv_group([
            (f[(0, 1)][0], xs[1], {}),        # 1x1 96->48 at 1/8
            (f[(0, 2)][0], xs[2], {}),        # 1x1 192->48 at 1/16
            (f[(1, 2)][0], xs[2], {}),        # 1x1 192->96 at 1/16
            (f[(2, 0)][0], xs[0], {}),        # 3x3 s2 48->48 (+ReLU), first hop of the 0->2 chain
            (f[(2, 1)][0], xs[1], {}),        # 3x3 s2 96->192
        ])
        y1, y2 = r.conv_group([
            (f[(1, 0)][0], xs[0], dict(add0=xs[1], add1=t12, add1_shift=1, relu=True)),
            (f[(2, 0)][1], u20, dict(add0=xs[2], add1=t21, relu=True)),
        ])
        # branch 0: identity + the two 1x1 terms evaluated at their own resolution, one HBM-bound pass
        return [r.upsum(xs[0], t01, 1, t02, 2), y1, y2]

    def run(self, r, x):
        """x: fp32 NCHW [S,3,H,W] on the device -> list of fp16 NHWC branch maps after stage3."""
        h = r.stem(x, self.stem_w, self.stem_scale, self.stem_bias, 64)
        # with I2R_HALO_CHAIN=1 consecutive halo-kernel launches (layer1's 13 convs, the 8 BasicBlock layers of every
        # module) run as chained grids; everything else (stride-2 / resampling problems, upsum) flushes the open chain
        # first.  Default: the context is inert and every layer is its own launch.
        with r.chain():
            return self._run_from_stem(r, h)

    def _run_from_stem(self, r, h):
        h = r.conv(self.conv2, h)
        for u in self.layer1:
            h = self._bottleneck(r, u, h)
        xs = []
        for t in self.trans1:
            y = h
            for L in (t or []):
                y = r.conv(L, y)
            xs.append(y)
        for mod in self.stage2:
            xs = self._fuse(r, mod, self._branches(r, mod, xs))
        nxt = []
        for i, t in enumerate(self.trans2):
            if t is None:
                nxt.append(xs[i])
            else:
                y = xs[-1]
                for L in t:
                    y = r.conv(L, y)
                nxt.append(y)
        xs = nxt
        for mod in self.stage3:
            xs = self._fuse(r, mod, self._branches(r, mod, xs))
        return xs
