"""sys.path setup shared by tests, bench.py and __graft_entry__.py.

The package directory name contains hyphens (it is fixed by the task layout), so it is put on
sys.path rather than imported by name: `i2r_b200` (engine) and, from its `lib/` sub-directory,
`models`, `config`, `utils` -- the same top-level module names the reference's tools/_init_paths.py
exposes, so `eval('models.' + cfg.MODEL.NAME + '.get_pose_net')` works unchanged.
"""
import os
import sys

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "intra-and-inter-human-relation-network-for-mpee_b200")
LIB = os.path.join(PKG, "lib")


def setup():
    for p in (LIB, PKG, REPO):
        if p not in sys.path:
            sys.path.insert(0, p)
    return PKG


setup()
