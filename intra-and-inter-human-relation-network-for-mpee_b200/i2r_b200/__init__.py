"""i2r_b200 -- host side of the B200-native I2R-Net forward path.

Layout:
  build.py    nvcc build of csrc/*.cu -> libi2r_sm100.so (in-tree, sm_100a only)
  capi.py     ctypes binding of include/i2r.h (fails loudly when the library is missing)
  packing.py  BatchNorm folding + weight packing into the tcgen05 core-matrix layout
  ops.py      thin op wrappers that fill i2r_conv_problem structs and launch on torch's stream
  config.py   yacs-compatible CfgNode + the reference's MODEL defaults
  synth.py    deterministic synthetic weights / inputs (tests, smoke, bench)
The reference-facing surface (get_pose_net / forward) lives in ../lib/models, mirroring the
reference's lib/models module names.
"""
from .version import __version__  # noqa: F401
