"""HRFormer-B first stage ("hrformer"): the drop-in SURFACE of the reference module of the same name
(lib/models/hrformer.py:2470-2533) -- same factory signature `get_pose_net(cfg, is_train, model_path, e2e_flag)`,
same hard-coded architecture (:2489-2525: stem, 2 Bottlenecks, three transformer stages of 1 / 4 / 2 modules with
78 / 156 / 312 / 624 channels, 2 / 4 / 8 / 16 heads, 7x7 windows, MLP ratio 4, zero-deconv 1x1 head) and the same
state_dict keys / shapes / dtypes (2074 backbone tensors + `keypoint_head.final_layer.*`; checked against the key
list dumped from the real reference, tests/golden/state_dict_interformer_coco_hrt_*.json), so the reference's
HRFormer checkpoints load with strict=True and `models.interformer.get_pose_net` resolves this first stage by name
exactly as interformer.py:139 does.

The forward runs as the device program of `i2r_b200/hrformer_program.py` (window attention over 49-token windows incl.
the zero-padded tokens and WITHOUT the relative position bias (:866-888), LayerNorm eps 1e-6, MlpDWBN = 1x1 GEMMs +
depthwise 3x3 + erf-GELU, bilinear fuse, 78/156/312/624 channels zero-padded to multiples of 16); the pinned oracle is
`oracle/i2r_oracle.hrformer_first_stage`.  There is deliberately no PyTorch or CPU fallback.
"""
import logging
import os

import torch
import torch.nn as nn

from i2r_b200 import capi

logger = logging.getLogger(__name__)

BN_MOMENTUM = 0.1

# lib/models/hrformer.py:2489-2525 (hard-coded in the reference's get_pose_net)
HRT_BASE = dict(
    stage1=dict(num_modules=1, num_branches=1, num_blocks=(2,), num_channels=(64,)),
    stage2=dict(num_modules=1, num_branches=2, num_blocks=(2, 2), num_channels=(78, 156), num_heads=(2, 4),
                num_mlp_ratios=(4, 4), num_window_sizes=(7, 7)),
    stage3=dict(num_modules=4, num_branches=3, num_blocks=(2, 2, 2), num_channels=(78, 156, 312),
                num_heads=(2, 4, 8), num_mlp_ratios=(4, 4, 4), num_window_sizes=(7, 7, 7)),
    stage4=dict(num_modules=2, num_branches=4, num_blocks=(2, 2, 2, 2), num_channels=(78, 156, 312, 624),
                num_heads=(2, 4, 8, 16), num_mlp_ratios=(4, 4, 4, 4), num_window_sizes=(7, 7, 7, 7)),
)


def _bn(c):
    return nn.BatchNorm2d(c, momentum=BN_MOMENTUM)


class _Bottleneck(nn.Module):          # reference :1244-1348 (expansion 4)
    def __init__(self, inplanes, planes, downsample):
        super().__init__()
        if downsample:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, bias=False), _bn(planes * 4))
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = _bn(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = _bn(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = _bn(planes * 4)


class _WindowMHA(nn.Module):           # MHA_ (:590-935): separate biased q/k/v/out projections + (unused) RPE table
    def __init__(self, dim, heads, ws):
        super().__init__()
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) * (2 * ws - 1), heads))
        coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
        rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += ws - 1
        rel[:, :, 1] += ws - 1
        rel[:, :, 0] *= 2 * ws - 1
        self.register_buffer("relative_position_index", rel.sum(-1))
        self.k_proj = nn.Linear(dim, dim)
        self.v_proj = nn.Linear(dim, dim)
        self.q_proj = nn.Linear(dim, dim)
        self.out_proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)


class _InterlacedPoolAttention(nn.Module):     # :1138-1180
    def __init__(self, dim, heads, ws):
        super().__init__()
        self.attn = _WindowMHA(dim, heads, ws)


class _MlpDWBN(nn.Module):             # :1044-1136
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Conv2d(dim, hidden, 1)
        self.norm1 = _bn(hidden)
        self.dw3x3 = nn.Conv2d(hidden, hidden, 3, 1, 1, groups=hidden)
        self.norm2 = _bn(hidden)
        self.fc2 = nn.Conv2d(hidden, dim, 1)
        self.norm3 = _bn(dim)


class _TransformerBlock(nn.Module):    # GeneralTransformerBlock (:1182-1240)
    def __init__(self, dim, heads, ws, mlp_ratio):
        super().__init__()
        self.attn = _InterlacedPoolAttention(dim, heads, ws)
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _MlpDWBN(dim, int(dim * mlp_ratio))


class _HRTModule(nn.Module):           # HighResolutionTransformerModule (:1454-1732)
    def __init__(self, cfg, multiscale_output):
        super().__init__()
        ch, nb = cfg["num_channels"], cfg["num_branches"]
        self.branches = nn.ModuleList(
            nn.Sequential(*[_TransformerBlock(ch[i], cfg["num_heads"][i], cfg["num_window_sizes"][i],
                                              cfg["num_mlp_ratios"][i]) for _ in range(cfg["num_blocks"][i])])
            for i in range(nb))
        fuse = []
        for i in range(nb if multiscale_output else 1):
            row = []
            for j in range(nb):
                if j > i:
                    row.append(nn.Sequential(nn.Conv2d(ch[j], ch[i], 1, bias=False), _bn(ch[i]),
                                             nn.Upsample(scale_factor=2 ** (j - i), mode="bilinear",
                                                         align_corners=False)))
                elif j == i:
                    row.append(None)
                else:
                    chain = []
                    for k in range(i - j):
                        cout = ch[i] if k == i - j - 1 else ch[j]
                        mods = [nn.Conv2d(ch[j], ch[j], 3, 2, 1, groups=ch[j], bias=False), _bn(ch[j]),
                                nn.Conv2d(ch[j], cout, 1, bias=False), _bn(cout)]
                        if k != i - j - 1:
                            mods.append(nn.ReLU(inplace=True))
                        chain.append(nn.Sequential(*mods))
                    row.append(nn.Sequential(*chain))
            fuse.append(nn.ModuleList(row))
        self.fuse_layers = nn.ModuleList(fuse)


def _transition(pre, cur):             # HRT._make_transition_layer (:1864-1918)
    layers = []
    for i in range(len(cur)):
        if i < len(pre):
            if cur[i] != pre[i]:
                layers.append(nn.Sequential(nn.Conv2d(pre[i], cur[i], 3, 1, 1, bias=False), _bn(cur[i]),
                                            nn.ReLU(inplace=True)))
            else:
                layers.append(None)
        else:
            chain = []
            for j in range(i + 1 - len(pre)):
                cout = cur[i] if j == i - len(pre) else pre[-1]
                chain.append(nn.Sequential(nn.Conv2d(pre[-1], cout, 3, 2, 1, bias=False), _bn(cout),
                                           nn.ReLU(inplace=True)))
            layers.append(nn.Sequential(*chain))
    return nn.ModuleList(layers)


class _HRT(nn.Module):                 # HRT (:1735-2100)
    def __init__(self, extra):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 3, 2, 1, bias=False)
        self.bn1 = _bn(64)
        self.conv2 = nn.Conv2d(64, 64, 3, 2, 1, bias=False)
        self.bn2 = _bn(64)
        nblk = extra["stage1"]["num_blocks"][0]
        self.layer1 = nn.Sequential(*[_Bottleneck(64 if i == 0 else 256, 64, downsample=(i == 0)) for i in range(nblk)])
        pre = [256]
        for si in (2, 3, 4):
            cfg = extra["stage%d" % si]
            setattr(self, "transition%d" % (si - 1), _transition(pre, list(cfg["num_channels"])))
            nmod = cfg["num_modules"]
            setattr(self, "stage%d" % si, nn.Sequential(*[
                _HRTModule(cfg, multiscale_output=not (si == 4 and mi == nmod - 1)) for mi in range(nmod)]))
            pre = list(cfg["num_channels"])


class _Head(nn.Module):                # TopDownSimpleHead with num_deconv_layers=0 (:2215-2348)
    def __init__(self, cin, cout):
        super().__init__()
        self.final_layer = nn.Conv2d(cin, cout, 1)


class HRFormer(nn.Module):
    def __init__(self, hrt_extra, head_in_channel, head_out_channel, num_deconv_layers):
        super().__init__()
        if num_deconv_layers != 0:
            raise NotImplementedError("HRFormer head with deconv layers (the reference builds it with 0, :2527)")
        self.backbone = _HRT(hrt_extra)
        self.keypoint_head = _Head(head_in_channel, head_out_channel)
        self.hrt_extra = hrt_extra
        # 'split' = split-operand GEMMs (fp16 hi+lo pairs); 'fp16' = single-pass kernels
        self.precision = os.environ.get("I2R_PRECISION_HRT", "split")
        self._program = None
        self._runner = None

    def build_program(self, device):
        """Fold BN, pad channels to multiples of 16, pack weights, upload: the device program the two-stage wrapper
        (or forward) runs."""
        from i2r_b200.hrformer_program import HRTProgram
        from i2r_b200.ops import channel_padding, split_precision
        sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        with split_precision(self.precision == "split"), channel_padding(16):
            return HRTProgram(self, sd, torch.device(device))

    def load_state_dict(self, *a, **kw):
        out = super().load_state_dict(*a, **kw)
        self._program = None
        return out

    def forward(self, x):
        """x fp32 [S,3,H,W] -> (branch-0 feature [S,78,H/4,W/4] fp32, heatmaps [S,K,H/4,W/4]) like the reference (:2477)."""
        from i2r_b200.ops import Runner
        from i2r_b200.packing import merge_pair
        dev = self.keypoint_head.final_layer.weight.device
        if dev.type != "cuda":
            raise capi.I2RError("hrformer forward runs on a CUDA (sm_100a) device only -- there is no CPU fallback")
        if self._program is None or self._program.device != dev:
            self._program = self.build_program(dev)
            self._runner = Runner(dev, 0)
            self._runner.split = self.precision == "split"
        with torch.no_grad():
            feat, heat = self._program.run(self._runner, x.to(dev, dtype=torch.float32).contiguous())
        f = merge_pair(feat) if self.precision == "split" else feat.float()
        return f[..., :78].permute(0, 3, 1, 2).contiguous(), heat


def get_pose_net(cfg, is_train, model_path="", e2e_flag=False):
    model = HRFormer(HRT_BASE, 78, cfg.MODEL.NUM_JOINTS, 0)
    if is_train:
        ckpt = torch.load(model_path, map_location="cpu")
        model.load_state_dict(ckpt.get("state_dict", ckpt), strict=False)
    logger.info("=> loading hrformer pretrained model {}".format(model_path))
    return model
