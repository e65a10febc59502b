"""Stage hand-off helpers: voxel_grid_to_points against the live reference's output, the .npz / camera
JSON round trips the notebooks perform between stages."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, DATA, pkg
from helpers import sha


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "handoff_golden.npz"))


@pytest.fixture(scope="module")
def grid():
    return np.load(os.path.join(GOLDEN, "aligner_golden.npz"))["grid"]


def check(g, stride, pts, cols, shp):
    assert [str(pts.dtype), str(cols.dtype)] == list(g[f"s{stride}_dtypes"])
    assert len(pts) == int(g[f"s{stride}_n"]) and tuple(shp) == tuple(g[f"s{stride}_shape"])
    assert sha(pts) == str(g[f"s{stride}_pts_sha"]) and sha(cols) == str(g[f"s{stride}_cols_sha"])


@pytest.mark.parametrize("stride", [1, 2, 3])
def test_oracle_voxel_grid_to_points(oracle, g, grid, stride):
    check(g, stride, *oracle.voxel_grid_to_points(grid, stride=stride))


@pytest.mark.gpu
@pytest.mark.parametrize("stride", [1, 2, 3])
def test_voxel_grid_to_points_gpu(g, grid, stride):
    vu = pkg("utils.voxel_utils")
    check(g, stride, *vu.voxel_grid_to_points(grid, stride=stride))
    with pytest.raises(NotImplementedError):
        vu.voxel_grid_to_points(grid[..., 0])


def test_io_round_trips(tmp_path, grid):
    """The loader must be importable and usable without a GPU: plain host I/O."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("io_utils", os.path.join(os.path.dirname(GOLDEN), "..",
                                                  "part-based-3d-reconstruction_b200", "utils", "io_utils.py"))
    io = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(io)
    p = io.save_voxel_grid(tmp_path / "stage1" / "Taj_voxel_grid.npz", grid)
    assert np.array_equal(io.load_voxel_grid(p), grid)
    assert np.array_equal(np.load(p)["voxel_grid"], grid)                # what nb2 cell 3 does
    with pytest.raises(ValueError):
        io.save_voxel_grid(tmp_path / "x.npz", grid[..., 0])
    ref_json = os.path.join(DATA, "results", "2.Perspective_Camera_Estimation", "Taj_camera_params_final.json")
    cams32 = io.load_camera_params(ref_json)
    assert cams32["front"]["cam_pos"].dtype == np.float32 and isinstance(cams32["front"]["f"], float)
    cams64 = io.load_camera_params(ref_json, np.float64)
    out = io.save_camera_params(tmp_path / "cams" / "Taj_camera_params_final.json", cams64)
    assert json.load(open(out)) == json.load(open(ref_json))             # lossless through float64
    assert io.to_json_safe({"a": (np.float32(1.5), np.arange(2))}) == {"a": [1.5, [0, 1]]}


@pytest.mark.gpu
def test_voxel_grid_to_points_edge_cases(oracle):
    vu = pkg("utils.voxel_utils")
    empty = np.zeros((5, 4, 6, 3), np.uint8)
    pts, cols, shp = vu.voxel_grid_to_points(empty, stride=2)
    assert pts.shape == (0, 3) and cols.shape == (0, 3) and shp == (4, 5, 6)
    rng = np.random.default_rng(1)
    g = (rng.random((7, 9, 5, 3)) < 0.2).astype(np.uint8) * rng.integers(1, 255, (7, 9, 5, 3), dtype=np.uint8)
    for stride in (1, 2, 4, 16):                                           # 16 > every extent: only voxel (0,0,0) is sampled
        got, want = vu.voxel_grid_to_points(g, stride=stride), oracle.voxel_grid_to_points(g, stride=stride)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and got[2] == want[2]
