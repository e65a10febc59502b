"""Depth-buffer visibility evaluator (utils/eval_helpers_intra.py:134-190) against vectors recorded from
the live reference: oracle on CPU, CUDA kernels on the GPU."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, pkg
from helpers import unpack

H, W = 139, 256


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "depth_golden.npz"))


@pytest.fixture(scope="module")
def grid():
    return np.load(os.path.join(GOLDEN, "aligner_golden.npz"))["grid"]


def cam_of(row, dt):
    return {"cam_pos": row[0:3].astype(dt), "target": row[3:6].astype(dt), "f": float(row[6]), "cx": float(row[7]),
            "cy": float(row[8])}


CASES = [(v, d) for v in ("front", "drone") for d in (np.float32, np.float64)]


@pytest.mark.parametrize("view,dt", CASES)
def test_oracle_matches_reference(oracle, g, grid, view, dt):
    key = f"{view}_{np.dtype(dt).name}"
    cam = cam_of(g[key + "_cam"], dt)
    z = oracle.compute_global_depth_buffer(grid, cam, H, W)
    assert z.dtype == np.float32 and np.array_equal(z, g[key + "_zbuf"])
    for tag, parts in (("min", ["front_minarets", "back_minarets"]), ("dome", ["dome"])):
        pts, _ = oracle.get_voxel_points_by_parts(grid, oracle.PART_COLORS, parts)
        assert np.array_equal(oracle.project_part_visible(pts, cam, z, H, W), unpack(g[f"{key}_{tag}_visible"], (H, W)).astype(bool))
        assert np.array_equal(oracle.project_part_visible(pts, cam, z, H, W, eps=0.75), unpack(g[f"{key}_{tag}_loose"], (H, W)).astype(bool))


@pytest.mark.gpu
@pytest.mark.parametrize("view,dt", CASES)
def test_cuda_matches_reference(oracle, g, grid, view, dt):
    eh = pkg("utils.eval_helpers_intra")
    key = f"{view}_{np.dtype(dt).name}"
    cam = cam_of(g[key + "_cam"], dt)
    z = eh.compute_global_depth_buffer(grid, cam, H, W)
    assert z.dtype == np.float32 and z.shape == (H, W) and np.array_equal(z, g[key + "_zbuf"])
    for tag, parts in (("min", ["front_minarets", "back_minarets"]), ("dome", ["dome"])):
        pts, _ = oracle.get_voxel_points_by_parts(grid, oracle.PART_COLORS, parts)
        got = eh.project_part_visible(pts, cam, z, H, W)
        assert got.dtype == bool and np.array_equal(got, unpack(g[f"{key}_{tag}_visible"], (H, W)).astype(bool))
        assert np.array_equal(eh.project_part_visible(pts, cam, z, H, W, eps=0.75), unpack(g[f"{key}_{tag}_loose"], (H, W)).astype(bool))
    a = unpack(g[f"{key}_min_visible"], (H, W)).astype(bool)
    assert eh._iou_bool(a, a) == 1.0 and np.isnan(eh._iou_bool(a & False, a & False))
