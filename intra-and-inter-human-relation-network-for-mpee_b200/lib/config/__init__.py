"""`from config import cfg, update_config` -- same import surface as the reference's lib/config."""
from .default import _C as cfg  # noqa: F401
from .default import update_config  # noqa: F401
