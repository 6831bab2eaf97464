"""Global config node `_C` and update_config, backed by i2r_b200.config (reference: lib/config/default.py)."""
from i2r_b200.config import cfg as _C  # noqa: F401
from i2r_b200.config import update_config  # noqa: F401
