"""Generate tests/golden/hrnet_c1.npz and state_dict_hrnet.json from the REAL reference `models.hrnet`
(lib/models/hrnet.py:275-487; build container only):  python tests/golden/make_golden_hrnet.py"""
import argparse
import copy
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


def main():
    cfg, update_config, models = ref_harness.import_reference()
    cfg = copy.deepcopy(cfg)
    yaml_rel = "coco/interformer_coco_w48_pure_en6.yaml"
    update_config(cfg, argparse.Namespace(cfg=os.path.join(ref_harness.REF_ROOT, "experiments", yaml_rel), opts=[],
                                          modelDir="", logDir="", dataDir=""))
    model = models.hrnet.get_pose_net(cfg, is_train=False).eval()
    model.load_state_dict(synth_state_dict(model.state_dict(), seed=0), strict=True)
    keys = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items()}
    with open(os.path.join(HERE, "state_dict_hrnet.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)
    x, _ = synth_inputs(1, 256, 192, seed=1)
    with torch.no_grad():
        out = model(x)
    np.savez_compressed(os.path.join(HERE, "hrnet_c1.npz"), out=out.numpy())
    print(tuple(out.shape), float(out.abs().max()), len(keys))


if __name__ == "__main__":
    main()
