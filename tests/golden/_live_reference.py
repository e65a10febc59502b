"""Import the LIVE reference (read-only checkout) with its visualisation deps stubbed.

Only usable in the build container, where ``/root/reference`` exists.  Used by
``make_golden.py`` to (re)generate the committed fixtures and by nothing else:
tests, smoke() and bench.py never import this.
"""
import os
import sys
from unittest import mock

REF_ROOT = os.environ.get("P3D_REFERENCE_ROOT", "/root/reference")


def load():
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"live reference not found at {REF_ROOT}")
    for n in ["matplotlib", "matplotlib.pyplot", "plotly", "plotly.graph_objects", "trimesh",
              "skimage", "skimage.measure", "ipywidgets", "IPython", "IPython.display"]:
        sys.modules.setdefault(n, mock.MagicMock(name=n))
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import utils.voxel_carving_utils as vc
    vc.tqdm = lambda it, **k: it
    import utils.camera_estimation as ce
    import utils.camera_geometry as cg
    import utils.config as cfg
    import utils.mask_utils as mu
    import utils.projection_utils as pu
    import utils.voxel_utils as vu

    class Ref:
        pass

    r = Ref()
    r.vc, r.ce, r.cg, r.cfg, r.mu, r.pu, r.vu = vc, ce, cg, cfg, mu, pu, vu
    r.root = REF_ROOT
    return r
