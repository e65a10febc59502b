"""Loader of the UNMODIFIED reference functions for bench.py's CPU arms.

`tools/install_ref.py` copies the reference's `utils/*.py` verbatim from /root/reference into `baseline/_ref/utils/`
(git-ignored, shipped to the GPU box by gpurun).  This module imports them from there with the plotting / widget modules
the image lacks replaced by `unittest.mock.MagicMock` (none of them is touched by the timed functions), and exposes the
six functions of the scoring path.  Nothing of the product package or of oracle/ is imported here.
"""
from __future__ import annotations

import importlib
import os
import sys
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")
_STUBS = ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "plotly", "plotly.graph_objects", "plotly.subplots",
          "trimesh", "skimage", "skimage.measure", "ipywidgets", "IPython", "IPython.display", "open3d"]
FILES = ["__init__.py", "config.py", "mask_utils.py", "voxel_utils.py", "voxel_carving_utils.py", "camera_geometry.py",
         "camera_estimation.py", "deformation_estimation.py", "projection_utils.py", "visualization.py"]


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_ROOT, "utils", f)) for f in FILES)


class Reference:
    """Namespace of the reference's own functions (utils/*.py under baseline/_ref)."""

    def __init__(self):
        if not available():
            raise ImportError(f"{REF_ROOT}/utils is missing: run `python tools/install_ref.py` where /root/reference exists")
        for n in _STUBS:
            if n not in sys.modules:
                try:
                    importlib.import_module(n)
                except Exception:
                    sys.modules[n] = mock.MagicMock(name=n)
        sys.dont_write_bytecode = True
        if "utils" in sys.modules and not getattr(sys.modules["utils"], "__file__", "").startswith(REF_ROOT):
            raise ImportError("another top-level package named `utils` is already imported")
        sys.path.insert(0, REF_ROOT)
        try:
            self.config = importlib.import_module("utils.config")
            self.mask_utils = importlib.import_module("utils.mask_utils")
            self.voxel_utils = importlib.import_module("utils.voxel_utils")
            self.camera_geometry = importlib.import_module("utils.camera_geometry")
            self.projection_utils = importlib.import_module("utils.projection_utils")
            self.camera_estimation = importlib.import_module("utils.camera_estimation")
            self.voxel_carving_utils = importlib.import_module("utils.voxel_carving_utils")
        finally:
            sys.path.remove(REF_ROOT)
        self.voxel_carving_utils.tqdm = lambda it, **k: it                     # silence the progress bar
        self.project_colored_voxels = self.projection_utils.project_colored_voxels   # projection_utils.py:5-23
        self.compute_partwise_iou = self.camera_estimation.compute_partwise_iou      # camera_estimation.py:770-787
        self.mask_parts_from_image = self.mask_utils.mask_parts_from_image           # mask_utils.py:89-97
        self.PART_COLORS = self.config.PART_COLORS

    def evaluate(self, pts, cols, seg_img, selected_labels, row, H, W):
        """The `evaluate` closure of launch_smart_aligner (camera_estimation.py:597-603) for one candidate row
        [cam_pos, target, f, cx, cy]: returns (per-part IoU dict, mean IoU)."""
        proj = self.project_colored_voxels(pts, cols, row[0:3], row[3:6], row[6], row[7], row[8], H, W)
        return self.compute_partwise_iou(proj, seg_img, selected_labels)

    def counts(self, pts, cols, seg_img, selected_labels, row, H, W):
        """(inter, union) integer counts per part as compute_partwise_iou forms them (:777-781), plus the mean IoU."""
        import numpy as np
        proj = self.project_colored_voxels(pts, cols, row[0:3], row[3:6], row[6], row[7], row[8], H, W)
        a, b = proj.reshape(-1, 3), seg_img.reshape(-1, 3)
        out = []
        for colour in selected_labels.values():
            pa, pb = np.all(a == colour, axis=1), np.all(b == colour, axis=1)
            out.append((int(np.logical_and(pa, pb).sum()), int(np.logical_or(pa, pb).sum())))
        _, mean_iou = self.compute_partwise_iou(proj, seg_img, selected_labels)
        return out, float(mean_iou)
