import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(GOLDEN, "data")
PKG = "part-based-3d-reconstruction_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pkg(sub=""):
    """Import a module of the product package (its directory name is not a Python identifier)."""
    return importlib.import_module(PKG + (("." + sub) if sub else ""))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def camera_golden():
    return np.load(os.path.join(GOLDEN, "camera_golden.npz"))


@pytest.fixture(scope="session")
def carve_golden():
    return np.load(os.path.join(GOLDEN, "carve_golden.npz"))


@pytest.fixture(scope="session")
def taj():
    """Config 2 inputs: the reference's stored Taj grid, its final camera JSON and both masks."""
    import cv2
    grid = np.load(os.path.join(DATA, "results", "1.Orthographic_Voxel_Carving", "Taj_voxel_grid.npz"))["voxel_grid"]
    cams = json.load(open(os.path.join(DATA, "results", "2.Perspective_Camera_Estimation", "Taj_camera_params_final.json")))

    def load(view, max_dim=None):                      # utils/mask_utils.py:14-33
        m = cv2.cvtColor(cv2.imread(os.path.join(DATA, "Taj", "masks", f"Taj_{view}_mask.png")), cv2.COLOR_BGR2RGB)
        if max_dim is not None:
            h, w = m.shape[:2]
            s = max_dim / max(h, w)
            m = cv2.resize(m, (int(w * s), int(h * s)), interpolation=cv2.INTER_NEAREST)
        return m

    return {"grid": grid, "cams": cams, "front": load("front", int(max(grid.shape))), "drone": load("drone")}
