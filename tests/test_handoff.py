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
    try:
        import matplotlib  # noqa: F401
    except ImportError:                                      # the scalar branch calls matplotlib's colormap, like the reference
        with pytest.raises(ImportError):
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


@pytest.mark.gpu
def test_extract_top_k_components_matches_scipy(oracle):
    """voxel_utils.py:22-31 on the device (26-connected labelling, extents from the component boxes) against the
    scipy-based restatement: touching-by-corner blobs (one component under 26-connectivity, several under 6), ties in
    height, k larger than the number of components, an absent colour."""
    import torch
    vu = pkg("utils.voxel_utils")
    rng = np.random.default_rng(12)
    colour, other = (0, 0, 255), (190, 0, 255)
    for shape, p in (((20, 31, 18), 0.18), ((9, 40, 12), 0.3), ((16, 16, 16), 0.08)):
        grid = np.zeros(shape + (3,), np.uint8)
        grid[rng.random(shape) < p] = colour
        grid[rng.random(shape) < 0.1] = other
        for k in (1, 2, 4, 1000):
            want = oracle.extract_top_k_components(grid, colour, k)
            got = vu.extract_top_k_components(grid, colour, k)
            assert got.dtype == np.uint8 and np.array_equal(got, want), (shape, k)
        t = vu.extract_top_k_components(torch.from_numpy(grid).cuda(), colour, 2)
        assert t.is_cuda and np.array_equal(t.cpu().numpy(), oracle.extract_top_k_components(grid, colour, 2))
        assert np.array_equal(vu.extract_top_k_components(grid, (1, 2, 3), 2), grid)
    diag = np.zeros((4, 4, 4, 3), np.uint8)
    diag[0, 0, 0] = diag[1, 1, 1] = diag[3, 3, 3] = colour                  # corner contact: {(0,0,0),(1,1,1)} is ONE component
    assert np.array_equal(vu.extract_top_k_components(diag, colour, 1), oracle.extract_top_k_components(diag, colour, 1))


@pytest.mark.gpu
def test_scalar_grid_to_points_branch(oracle, monkeypatch):
    """voxel_utils.py:47-49: the scalar branch.  matplotlib is not in this image, so a stand-in colormap (a fixed
    function of the normalised value) is injected; what is checked is everything the reference computes around the
    colormap call -- points, their order, the normalised values with the reference's index/extent pairing, the shape
    tuple -- against the NumPy restatement."""
    import sys
    import types
    vu = pkg("utils.voxel_utils")
    seen = {}

    def get_cmap(name):
        def cmap(vals):
            seen["name"], seen["vals"] = name, np.asarray(vals)
            v = np.asarray(vals, dtype=np.float64)
            return np.stack([v, 1.0 - v, 0.5 * v, np.ones_like(v)], axis=1)
        return cmap

    plt = types.ModuleType("matplotlib.pyplot")
    plt.get_cmap = get_cmap
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    monkeypatch.setitem(sys.modules, "matplotlib", mpl)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    rng = np.random.default_rng(2)
    grid = (rng.random((13, 10, 17)) < 0.3).astype(np.uint8) * 7
    for axis, stride in (("z", 2), ("x", 1), ("y", 3)):
        pts, cols, shp = vu.voxel_grid_to_points(grid, axis=axis, colormap="magma", stride=stride)
        wp, wv, ws = oracle.scalar_grid_points(grid, axis, stride)
        assert pts.dtype == np.float32 and np.array_equal(pts, wp) and shp == ws
        assert seen["name"] == "magma" and np.array_equal(seen["vals"], wv)
        assert np.array_equal(cols, (get_cmap("magma")(wv)[:, :3] * 255).astype(np.uint8))


def test_top_k_oracle_pinned_on_the_live_reference(oracle):
    """tests/golden/topk_golden.npz (live reference, make_golden.topk_golden) pins the oracle's restatement of
    extract_top_k_components; the GPU test above compares the device path with that restatement."""
    g = np.load(os.path.join(GOLDEN, "topk_golden.npz"))
    colour = tuple(int(v) for v in g["colour"])
    for i in range(int(g["n"])):
        for k in (1, 2, 4):
            assert np.array_equal(oracle.extract_top_k_components(g[f"g{i}"], colour, k), g[f"g{i}_k{k}"]), (i, k)


@pytest.mark.gpu
def test_top_k_device_matches_the_live_reference():
    vu = pkg("utils.voxel_utils")
    g = np.load(os.path.join(GOLDEN, "topk_golden.npz"))
    colour = tuple(int(v) for v in g["colour"])
    for i in range(int(g["n"])):
        for k in (1, 2, 4):
            assert np.array_equal(vu.extract_top_k_components(g[f"g{i}"], colour, k), g[f"g{i}_k{k}"]), (i, k)
