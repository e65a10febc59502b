"""Camera initialisation chain of notebook 2 (utils/camera_estimation.py:20-344) against vectors recorded
from the live reference (tests/golden/make_golden.py init; its two library shims are described there): the oracle on
the CPU, the CUDA-backed package functions on the GPU."""
import contextlib
import io
import os

import numpy as np
import pytest

from conftest import GOLDEN, pkg
from helpers import sha

PARTS = ["front_minarets", "back_minarets"]


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "init_golden.npz"))


@pytest.fixture(scope="module")
def scene():
    ag = np.load(os.path.join(GOLDEN, "aligner_golden.npz"))
    return ag["grid"], ag["image"]


def row_of(p):
    return np.array([*p["cam_pos"], *p["target"], p["f"], p["cx"], p["cy"]], dtype=np.float64)


def check_init(g, init):
    assert np.array_equal(row_of(init), g["init_row"])
    assert [str(np.asarray(init[k]).dtype) for k in ("cam_pos", "target", "f", "cx", "cy")] == list(g["init_dtypes"])


def check_parts(g, vparts, mparts, shape_hw):
    assert sorted(vparts) == ["LM1", "LM2", "RM1", "RM2"]
    for k, v in vparts.items():
        assert len(v) == int(g[f"vox_{k}_n"]) and sha(np.asarray(v).astype(np.int64)) == str(g[f"vox_{k}_sha"]), k
    assert list(mparts.keys()) == list(g["mask_keys"])
    n = shape_hw[0] * shape_hw[1]
    for k, m in mparts.items():
        assert m.dtype == np.uint8 and np.array_equal(m.astype(bool).ravel(), np.unpackbits(g[f"mask_{k}"])[:n].astype(bool)), k


def check_kps(g, vk, ik):
    assert list(vk.keys()) == list(g["vk_keys"]) and list(ik.keys()) == list(g["ik_keys"])
    assert np.array_equal(np.array([vk[k] for k in vk], dtype=np.float64), g["vk"])
    assert np.array_equal(np.array([ik[k] for k in ik], dtype=np.float64), g["ik"])


def golden_selection(g):
    vsel = {str(k): v for k, v in zip(g["sel_keys"], g["sel_v"])}
    isel = {str(k): tuple(v) for k, v in zip(g["sel_keys"], g["sel_i"])}
    return vsel, isel


def init_of(g):
    r = g["init_row"]
    return {"cam_pos": r[0:3].copy(), "target": r[3:6].astype(np.float32), "f": r[6], "cx": r[7], "cy": r[8]}


# ---- CPU: oracle vs the live reference ------------------------------------------------------------------------------
def test_oracle_init_chain(oracle, g, scene):
    grid, image = scene
    colours = [oracle.PART_COLORS[p] for p in PARTS]
    init, _ = oracle.initial_params_matching_bbox(grid, image, oracle.PART_COLORS, PARTS)
    check_init(g, init)
    vparts, mparts = oracle.minaret_voxels_by_label(grid, colours), oracle.minaret_masks_by_label(image, colours)
    check_parts(g, vparts, mparts, image.shape[:2])
    check_kps(g, oracle.top_bottom_voxel_points(vparts), oracle.top_bottom_image_points(mparts))
    vsel, isel = oracle.minaret_kps_for_view(grid, image, colours)
    gv, gi = golden_selection(g)
    assert set(vsel) == set(gv)
    for k in gv:
        assert np.array_equal(vsel[k], gv[k]) and tuple(isel[k]) == gi[k]


@pytest.mark.parametrize("loss", ["L2", "L1"])
def test_oracle_keypoint_fit(oracle, g, scene, loss):
    vsel, isel = golden_selection(g)
    fit, _ = oracle.optimize_camera_with_keypoints(vsel, isel, scene[1], init_of(g), loss_type=loss)
    assert np.array_equal(row_of(fit), g[f"fit_{loss}"])


# ---- GPU: the package vs the live reference ----------------------------------------------------------------------------
@pytest.mark.gpu
def test_init_chain_gpu(oracle, g, scene):
    ce = pkg("utils.camera_estimation")
    grid, image = scene
    colours = [oracle.PART_COLORS[p] for p in PARTS]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        init = ce.auto_compute_initial_params_matching_bbox(grid, image, oracle.PART_COLORS, PARTS)
    check_init(g, init)
    assert buf.getvalue() == str(g["init_log"])
    vparts, mparts = ce.extract_minaret_voxels_by_label(grid, colours), ce.extract_minaret_masks_by_label(image, colours)
    check_parts(g, vparts, mparts, image.shape[:2])
    check_kps(g, ce.extract_top_bottom_voxel_points(vparts), ce.extract_top_bottom_image_points(mparts))
    vsel, isel = ce.extract_minaret_kps_for_view(grid, image, colours)
    gv, gi = golden_selection(g)
    assert set(vsel) == set(gv)
    for k in gv:
        assert np.array_equal(vsel[k], gv[k]) and tuple(isel[k]) == gi[k]
    with pytest.raises(ValueError):
        ce.extract_minaret_voxels_by_label(grid, [(1, 2, 3)])
    with pytest.raises(ValueError):
        ce.extract_minaret_masks_by_label(image, [(1, 2, 3)])


@pytest.mark.gpu
@pytest.mark.parametrize("loss", ["L2", "L1"])
def test_keypoint_fit_gpu(g, scene, loss):
    ce = pkg("utils.camera_estimation")
    vsel, isel = golden_selection(g)
    fit = ce.optimize_camera_with_keypoints(vsel, isel, scene[1], init_of(g), loss_type=loss, verbose=False)
    assert np.array_equal(row_of(fit), g[f"fit_{loss}"])


@pytest.mark.gpu
def test_label8_matches_oracle_on_random_masks(oracle):
    ce = pkg("utils.camera_estimation")
    nv = pkg("utils._native")
    import torch
    rng = np.random.default_rng(3)
    for shape, p in (((37, 53), 0.45), ((128, 96), 0.6), ((5, 7), 0.3), ((64, 64), 0.0), ((33, 1), 0.7)):
        m = (rng.random(shape) < p).astype(np.uint8)
        want, n = oracle.label8_2d(m)
        labels, got_n, bbox, sums = ce._label_image8(torch.from_numpy(m).cuda())
        assert got_n == n and np.array_equal(labels.cpu().numpy()[0], want)
        for cid in range(1, n + 1):
            yy, xx = np.nonzero(want == cid)
            assert sums[cid - 1, 0] == len(yy) and sums[cid - 1, 2] == yy.sum() and sums[cid - 1, 3] == xx.sum()


@pytest.mark.gpu
def test_aligner_default_init_uses_bbox_init(g, scene, oracle):
    """launch_smart_aligner(init_params=None) starts from auto_compute_initial_params_matching_bbox (:525-526)."""
    ce = pkg("utils.camera_estimation")
    grid, image = scene
    with contextlib.redirect_stdout(io.StringIO()):
        saved = ce.launch_smart_aligner(grid, image, oracle.PART_COLORS, parts_for_alignment=PARTS)
    p = saved.aligner.get_params()
    assert np.allclose(row_of(p), g["init_row"])


@pytest.mark.gpu
def test_visualize_reprojection_prints_reference_table(g, scene):
    pu = pkg("utils.projection_utils")
    vsel, isel = golden_selection(g)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        pu.visualize_reprojection(scene[1], vsel, isel, init_of(g), title="Front | Initial Reprojection")
    assert buf.getvalue() == str(g["reproj_log"])
