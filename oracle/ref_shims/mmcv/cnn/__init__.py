"""mmcv.cnn stand-in: the builders return plain torch layers; the norm-layer name becomes a
state_dict key in the reference (hrformer.py:1269-1299), so it must be 'bn'+postfix."""
import torch.nn as nn


def build_conv_layer(cfg, *args, **kwargs):
    return nn.Conv2d(*args, **kwargs)


def build_norm_layer(cfg, num_features, postfix=""):
    kind = (cfg or {}).get("type", "BN")
    assert kind in ("BN", "SyncBN"), kind
    kw = {k: v for k, v in (cfg or {}).items() if k not in ("type", "requires_grad")}
    return "bn" + str(postfix), nn.BatchNorm2d(num_features, **kw)


def build_upsample_layer(cfg, *args, **kwargs):
    kind = cfg.get("type")
    assert kind == "deconv", kind
    return nn.ConvTranspose2d(*args, **kwargs)


def constant_init(module, val, bias=0):
    if getattr(module, "weight", None) is not None:
        nn.init.constant_(module.weight, val)
    if getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)


def kaiming_init(module, a=0, mode="fan_out", nonlinearity="relu", bias=0, distribution="normal"):
    if getattr(module, "weight", None) is not None:
        nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)


def normal_init(module, mean=0, std=1, bias=0):
    if getattr(module, "weight", None) is not None:
        nn.init.normal_(module.weight, mean, std)
    if getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)
