"""Generate tests/golden/*.npz and *.json from the REAL reference (run in the build container only):

    python tests/golden/make_golden.py

For each case: build the reference model from its own yaml, load the deterministic synthetic weights
(i2r_b200.synth.synth_state_dict -- keyed by parameter name), run the reference forward on the
synthetic inputs in fp32 on CPU, and store the heatmaps plus two intermediate taps (forward hooks on
`reduce` and `global_encoder`).  The key/shape list of the state_dict is stored as JSON so the
drop-in module's parameter surface can be checked without the reference.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "intra-and-inter-human-relation-network-for-mpee_b200"))

from oracle import ref_harness  # noqa: E402
from i2r_b200.synth import synth_inputs, synth_state_dict  # noqa: E402

CASES = {
    # name: (yaml, length, H, W)
    "vanilla_c1": ("coco/interformer_coco_w48_pure_en6.yaml", [1], 256, 192),
    "vanilla_ragged": ("coco/interformer_coco_w48_pure_en6.yaml", [2, 1], 256, 192),
    # two-stage, TransPose-H first stage (6 intra layers over 3072 tokens per crop), 4 inter layers, multiplex deconv
    "tph2stage_ragged": ("coco/interformer_coco_tph_192_p4_b4.yaml", [2, 1], 256, 192),
    # `interformer` naming: CrowdPose TransPose-H (4 intra / 2 inter layers, no multi-pos, distinct deconvs, 14 joints)
    "tph_crowdpose_ragged": ("crowdpose/interformer_crowdpose_tph_192_p6_b4.yaml", [1, 2], 256, 192),
    # HRFormer-B first stage (window attention incl. padded tokens, MlpDWBN, bilinear fuse) + 2 inter layers, d = 78
    "hrt2stage_ragged": ("coco/interformer_coco_hrt_192_p2_b12.yaml", [2, 1], 256, 192),
    # 384x288: 96x72 branch-0 maps padded to 98x77 windows, 24x18 token maps per person
    "hrt288_c1": ("coco/interformer_coco_hrt_288_p2_b4.yaml", [1], 384, 288),
    # BASELINE shapes of the HRFormer-B families, one image per rank: C4 = 8 persons at 256x192 (inter-human sequence of
    # 1536 tokens), C5 = 12 persons at 384x288 (5184 tokens); plus a three-image ragged batch.  Heatmaps are stored
    # subsampled (every 4th row / column, SUBSAMPLE below) to keep the fixtures small.
    "hrt_c4_rank": ("coco/interformer_coco_hrt_192_p2_b12.yaml", [8], 256, 192),
    "hrt_c5_rank": ("coco/interformer_coco_hrt_288_p2_b4.yaml", [12], 384, 288),
    "hrt_ragged3": ("coco/interformer_coco_hrt_192_p2_b12.yaml", [3, 1, 2], 256, 192),
    # compat extras: no first stage (stand-alone HRNet backbone, lib/models/backbone.py) + UpConv upsampling
    "hrnet_upconv_ragged": ("crowdpose/interformer_crowdpose_tph_192_p6_b4.yaml", [2, 1], 256, 192),
    # `res` multi-person position embedding (resnet18 stem on the box masks), the shipped OCHuman TransPose-H config
    "tph_ochuman_res_ragged": ("OCHuman/interformer_ochuman_tph_192_p3_b8.yaml", [1, 2], 256, 192),
}
OPTS = {"hrnet_upconv_ragged": ["MODEL.SINGLEFORMER", "", "MODEL.UPSAMPLE_TYPE", "upconv", "MODEL.INIT_WEIGHTS", False]}
SUBSAMPLE = {"hrt_c4_rank": 4, "hrt_c5_rank": 4, "hrt_ragged3": 4}


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    models = {}
    for name, (yaml_rel, length, h, w) in CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        mkey = (yaml_rel, name if name in OPTS else "")
        if mkey not in models:
            cfg, model = ref_harness.build_reference_model(yaml_rel, OPTS.get(name, ()))
            sd = synth_state_dict(model.state_dict(), seed=0)
            model.load_state_dict(sd, strict=True)
            models[mkey] = model
            keys = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items()}
            tag = cfg.MODEL.NAME if cfg.MODEL.NAME == "interformer_pureMulti" else os.path.basename(yaml_rel)[:-5]
            if name in OPTS:
                tag = name
            with open(os.path.join(HERE, "state_dict_%s.json" % tag), "w") as f:
                json.dump(keys, f, indent=0, sort_keys=True)
        model = models[mkey]
        x, pm = synth_inputs(sum(length), h, w, seed=1)
        taps = {}
        hooks = []
        only = [a for a in sys.argv[1:] if not a.startswith("-")]
        if hasattr(model, "reduce"):
            hooks.append(model.reduce.register_forward_hook(lambda m, i, o: taps.__setitem__("reduce", o.detach())))
        if hasattr(model, "global_encoder"):
            hooks.append(model.global_encoder.register_forward_hook(
                lambda m, i, o: taps.__setitem__("encoded_lbc", o.detach())))
        with torch.no_grad():
            out = model(x, pm, length)
        for hk in hooks:
            hk.remove()
        arrays = {"length": np.asarray(length, dtype=np.int64)}
        sub = SUBSAMPLE.get(name, 1)
        if sub > 1:
            arrays["subsample"] = np.asarray(sub, dtype=np.int64)
        if isinstance(out, dict):
            for k, v in out.items():
                arrays["out_" + k] = v.numpy()[:, :, ::sub, ::sub]
                arrays["absmax_" + k] = np.asarray(float(v.abs().max()))
        else:
            arrays["out"] = out.numpy()[:, :, ::sub, ::sub]
        for k, v in taps.items():
            arrays["tap_" + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        print(name, {k: (v.shape, float(np.abs(v).max())) for k, v in arrays.items()})


if __name__ == "__main__":
    main()
