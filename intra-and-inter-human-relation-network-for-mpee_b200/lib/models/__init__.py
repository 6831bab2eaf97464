"""Model registry: `eval('models.' + cfg.MODEL.NAME + '.get_pose_net')` must resolve exactly as in the
reference (tools/test.py:87), so the sub-module names are part of the API (lib/models/__init__.py:16-23)."""
import models.hrnet  # noqa: F401
import models.backbone  # noqa: F401
import models.interformer_pureMulti  # noqa: F401
import models.transpose_h  # noqa: F401
import models.interformer  # noqa: F401
import models.interformer_2stage  # noqa: F401
import models.hrformer  # noqa: F401
