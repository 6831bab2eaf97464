"""TEST INFRASTRUCTURE: import the REAL reference (read-only, /root/reference) on CPU through the
import shims in oracle/ref_shims (yacs / timm / mmcv are absent from this image).

Runs only in the build container (the GPU box has no /root/reference); its products are the committed
fixtures under tests/golden/.  Must not be imported together with this repo's own `lib/` (both expose
top-level `models` / `config` / `utils`), so callers run it in a separate process.
"""
import argparse
import os
import sys

REF_ROOT = os.environ.get("I2R_REF", "/root/reference")
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "lib", "models"))


def import_reference():
    """Returns (cfg, update_config, models) of the reference package."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    for p in (os.path.join(REF_ROOT, "lib"), SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    from config import cfg, update_config  # noqa: E402  (reference lib/config)
    import models  # noqa: E402             (reference lib/models)
    return cfg, update_config, models


def build_reference_model(yaml_rel, opts=()):
    """Construct the reference model of experiments/<yaml_rel> exactly as tools/test.py:87 does."""
    import copy
    cfg, update_config, models = import_reference()
    cfg = copy.deepcopy(cfg)
    args = argparse.Namespace(cfg=os.path.join(REF_ROOT, "experiments", yaml_rel), opts=list(opts), modelDir="",
                              logDir="", dataDir="")
    update_config(cfg, args)
    model = eval("models." + cfg.MODEL.NAME + ".get_pose_net")(cfg, is_train=False)
    return cfg, model.eval()
