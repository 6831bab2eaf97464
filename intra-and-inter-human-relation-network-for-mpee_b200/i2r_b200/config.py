"""Config surface: a yacs-compatible CfgNode (attribute + item access, yaml merge, KEY VAL overrides)
and the MODEL defaults the forward path reads (reference: lib/config/default.py:36-76).

When the real `yacs` package and the reference's lib/config are present they can be used instead --
the model factories only rely on attribute/item access.
"""
import ast
import copy
import os

import yaml


class CfgNode(dict):
    def __init__(self, init=None, new_allowed=False):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def defrost(self):
        return self

    def freeze(self):
        return self

    def clone(self):
        return copy.deepcopy(self)

    def merge_from_dict(self, other):
        for k, v in other.items():
            if isinstance(v, dict):
                if isinstance(self.get(k), CfgNode):
                    self[k].merge_from_dict(v)
                else:
                    self[k] = CfgNode(v)
            else:
                self[k] = list(v) if isinstance(v, tuple) else v

    def merge_from_file(self, path):
        with open(path, "r") as f:
            self.merge_from_dict(yaml.safe_load(f) or {})

    def merge_from_list(self, opts):
        if len(opts) % 2:
            raise ValueError("override list must be KEY VAL pairs")
        for key, val in zip(opts[0::2], opts[1::2]):
            node = self
            parts = key.split(".")
            for part in parts[:-1]:
                node = node[part]
            if isinstance(val, str):
                try:
                    val = ast.literal_eval(val)
                except (ValueError, SyntaxError):
                    pass
            node[parts[-1]] = val


def default_cfg():
    """Defaults for the keys the forward path reads ('Config / flags')."""
    c = CfgNode()
    c.OUTPUT_DIR = ""
    c.LOG_DIR = ""
    c.DATA_DIR = ""
    c.GPUS = (0,)
    c.MODEL = CfgNode(dict(
        NAME="interformer", SINGLEFORMER=None, SINGLE_MODEL="", NORMALIZE_BEFORE=False, END2END=False,
        BACKBONE_FIX=False, SINGLEFORMER_FIX=False, INIT_WEIGHTS=True, PRETRAINED="", NUM_JOINTS=17,
        IMAGE_SIZE=[256, 256], HEATMAP_SIZE=[64, 64], TRANS_SIZE=[16, 12], SIGMA=2, HRNET_RES_LAYER=0,
        EXTRA=CfgNode(), BOTTLENECK_NUM=0, DIM_MODEL=256, DIM_FEEDFORWARD=512, ENCODER_LAYERS=6,
        ENCODER_MULTI_LAYERS=4, USE_MULTI_POS=True, N_HEAD=8, ATTENTION_ACTIVATION="relu",
        POS_EMBEDDING="learnable", SINGLE_POS_EMBEDDING="sine", PE_ONLY_AT_BEGIN=False,
        INTER_SUPERVISION=True, UPSAMPLE_TYPE="multiplex", MULTI_POS_EMBEDDING="conv",
        ATTENTION_TYPE="default", WINDOW_SIZE=4, MULTI_POS_EMBEDDING_DIM=96, DOMAIN_TRANS=False,
    ))
    c.DATASET = CfgNode(dict(ROOT="", DATASET="coco"))
    c.TEST = CfgNode(dict(MODEL_FILE="", FLIP_TEST=False))
    return c


cfg = default_cfg()

EXPERIMENTS_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "experiments")


def update_config(cfg_node, args):
    """Same contract as the reference's update_config (lib/config/default.py:164-191)."""
    cfg_node.defrost()
    cfg_node.merge_from_file(args.cfg)
    cfg_node.merge_from_list(list(getattr(args, "opts", []) or []))
    if getattr(args, "modelDir", ""):
        cfg_node.OUTPUT_DIR = args.modelDir
    if getattr(args, "logDir", ""):
        cfg_node.LOG_DIR = args.logDir
    if getattr(args, "dataDir", ""):
        cfg_node.DATA_DIR = args.dataDir
    cfg_node.DATASET.ROOT = os.path.join(cfg_node.DATA_DIR, cfg_node.DATASET.ROOT)
    cfg_node.MODEL.PRETRAINED = os.path.join(cfg_node.DATA_DIR, cfg_node.MODEL.PRETRAINED)
    if cfg_node.TEST.MODEL_FILE:
        cfg_node.TEST.MODEL_FILE = os.path.join(cfg_node.DATA_DIR, cfg_node.TEST.MODEL_FILE)
    cfg_node.freeze()


def load_experiment(name, opts=()):
    """cfg for experiments/<name> (e.g. 'coco/interformer_coco_w48_pure_en6.yaml')."""
    c = default_cfg()
    path = name if os.path.isabs(name) else os.path.join(EXPERIMENTS_DIR, name)
    c.merge_from_file(path)
    c.merge_from_list(list(opts))
    return c
