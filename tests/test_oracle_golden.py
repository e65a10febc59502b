"""The oracle (oracle/) against the golden vectors generated from the LIVE reference
(tests/golden/make_golden.py).  CPU only: this is what pins the oracle."""
import io
import warnings

import numpy as np
import pytest
import scipy.ndimage

from helpers import EXTRUSION_DEPTHS, GROUP_JOBS, PART_SYMMETRY, row_to_args, sha, unpack


def test_look_at_matches_reference(oracle, camera_golden):
    g = camera_golden
    for e, t, R64, R32 in zip(g["lookat_eye"], g["lookat_target"], g["lookat_R64"], g["lookat_R32"]):
        assert np.array_equal(oracle.look_at_rotation(e, t), R64)
        r = oracle.look_at_rotation(e.astype(np.float32), t.astype(np.float32))
        assert r.dtype == np.float32 and np.array_equal(r, R32)


@pytest.mark.parametrize("view", ["front", "drone"])
@pytest.mark.parametrize("tag", ["min", "all"])
def test_taj_projection_and_iou(oracle, camera_golden, taj, view, tag):
    g = camera_golden
    parts = ["front_minarets", "back_minarets"] if tag == "min" else [p for p in oracle.PART_COLORS if p != "background"]
    img = taj[view]
    H, W = img.shape[:2]
    pts, cols = oracle.get_voxel_points_by_parts(taj["grid"], oracle.PART_COLORS, parts)
    assert pts.shape[0] == int(g[f"taj_{view}_{tag}_npts"]) and pts.dtype == np.float32
    seg = oracle.mask_parts_from_image(img, oracle.PART_COLORS, parts)
    sel = {p: oracle.PART_COLORS[p] for p in parts}
    cand = g[f"taj_{view}_{tag}_cand"]
    for k, row in enumerate(cand):
        proj = oracle.project_colored_voxels(pts, cols, *row_to_args(row), H, W)
        assert sha(proj) == str(g[f"taj_{view}_{tag}_sha"][k])
        if k == 0:
            assert np.array_equal(proj, g[f"taj_{view}_{tag}_image0"])
        inter, uni = oracle.partwise_counts(proj, seg, sel)
        assert np.array_equal(np.stack([inter, uni], 1), g[f"taj_{view}_{tag}_counts"][k])
        assert oracle.compute_partwise_iou(proj, seg, sel)[1] == g[f"taj_{view}_{tag}_scores"][k]


@pytest.mark.parametrize("view", ["front", "drone"])
def test_taj_projection_float32(oracle, camera_golden, taj, view):
    g = camera_golden
    parts = ["front_minarets", "back_minarets"]
    img = taj[view]
    H, W = img.shape[:2]
    pts, cols = oracle.get_voxel_points_by_parts(taj["grid"], oracle.PART_COLORS, parts)
    seg = oracle.mask_parts_from_image(img, oracle.PART_COLORS, parts)
    sel = {p: oracle.PART_COLORS[p] for p in parts}
    for k, row in enumerate(g[f"taj_{view}_f32_cand"]):
        proj = oracle.project_colored_voxels(pts, cols, *row_to_args(row, np.float32), H, W)
        assert sha(proj) == str(g[f"taj_{view}_f32_sha"][k])
        inter, uni = oracle.partwise_counts(proj, seg, sel)
        assert np.array_equal(np.stack([inter, uni], 1), g[f"taj_{view}_f32_counts"][k])


def test_adversarial_cameras(oracle, camera_golden):
    g = camera_golden
    parts = ["dome", "plinth", "windows"]
    pts, cols = oracle.get_voxel_points_by_parts(g["adv_grid"], oracle.PART_COLORS, parts)
    seg = oracle.mask_parts_from_image(g["adv_image"], oracle.PART_COLORS, parts)
    sel = {p: oracle.PART_COLORS[p] for p in parts}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k, row in enumerate(g["adv_cand"]):
            proj = oracle.project_colored_voxels(pts, cols, *row_to_args(row), 20, 24)
            assert sha(proj) == str(g["adv_sha"][k]), k
            inter, uni = oracle.partwise_counts(proj, seg, sel)
            assert np.array_equal(np.stack([inter, uni], 1), g["adv_counts"][k])
            assert oracle.compute_partwise_iou(proj, seg, sel)[1] == g["adv_scores"][k]


def test_affine_restatement_matches_scipy_vectors(oracle, carve_golden):
    g = carve_golden
    for k in range(int(g["aff_n"])):
        shape = tuple(g[f"aff{k}_shape"])
        vol = unpack(g[f"aff{k}_vol"], shape)
        M = oracle.rotation_matrix_inv(int(g[f"aff{k}_angle"]))
        ctr = np.array(shape) / 2
        got = oracle.affine_order1(vol, M, ctr - M @ ctr)
        assert np.array_equal(got, unpack(g[f"aff{k}_out"], shape)), (k, shape)
        # and against the scipy installed next to the tests (same dependency the reference calls)
        live = scipy.ndimage.affine_transform(vol, M, offset=ctr - M @ ctr, order=1, mode="constant", cval=0)
        assert np.array_equal(got, live)


def test_label_restatement_matches_scipy_vectors(oracle, carve_golden):
    g = carve_golden
    for i in range(int(g["lab_n"])):
        shape = tuple(g[f"lab{i}_shape"])
        m = unpack(g[f"lab{i}_mask"], shape)
        lab, n = oracle.label6(m)
        assert n == int(g[f"lab{i}_n"]) and np.array_equal(lab, g[f"lab{i}_out"])
    m = np.random.default_rng(3).random((17, 9, 22)) < 0.5
    view = np.flip(m.transpose(2, 1, 0), axis=1)
    a, n = scipy.ndimage.label(view)
    b, n2 = oracle.label6(view)
    assert n == n2 and np.array_equal(a, b)


@pytest.mark.parametrize("case", ["Bibi_64", "Taj_96", "Akbar_128", "Bibi_256"])
def test_real_mask_carving(oracle, carve_golden, case):
    g = carve_golden
    key = "real_" + case
    sem, ext, binm = g[key + "_sem"], g[key + "_ext"], g[key + "_bin"]
    grid = oracle.global_carve(binm, ext, 90)
    assert sha(grid) == str(g[key + "_global_sha"])
    assert np.count_nonzero(grid.any(-1)) == int(g[key + "_global_occ"])
    log = []
    out = oracle.partwise_carve(grid, ext, sem, oracle.PART_COLORS_NP, GROUP_JOBS, PART_SYMMETRY, EXTRUSION_DEPTHS, log=log)
    assert sha(out) == str(g[key + "_partwise_sha"])
    assert np.count_nonzero(out.any(-1)) == int(g[key + "_partwise_occ"])
    assert "\n".join(log) == str(g[key + "_log"]).strip("\n")
    if key + "_partwise" in g.files:
        assert np.array_equal(out, g[key + "_partwise"])


def test_synthetic_quirk_carving(oracle, carve_golden):
    g = carve_golden
    for tag in g["syn_cases"]:
        key = f"syn_{tag}"
        sem, ext, binm = g[key + "_sem"], g[key + "_ext"], g[key + "_bin"]
        grid = oracle.global_carve(binm, ext, 90)
        assert np.array_equal(grid, g[key + "_global"]), tag
        assert np.array_equal(oracle.part_carve(grid, ext, GROUP_JOBS), g[key + "_partcarve"]), tag
        out = oracle.partwise_carve(grid, ext, sem, oracle.PART_COLORS_NP, GROUP_JOBS, PART_SYMMETRY, EXTRUSION_DEPTHS)
        assert np.array_equal(out, g[key + "_partwise"]), tag
        out2 = oracle.partwise_carve(grid, ext, sem, oracle.PART_COLORS_NP, GROUP_JOBS, PART_SYMMETRY, EXTRUSION_DEPTHS,
                                     recolor_back_minarets=False)
        assert sha(out2) == str(g[key + "_partwise_norecolor_sha"]), tag


def test_per_part_scoring_core(oracle, camera_golden, taj):
    """visualize_voxel_projection_iou's scoring core (camera_estimation.py:381-403, 433-447)."""
    g = camera_golden
    img = taj["front"]
    H, W = img.shape[:2]
    p = taj["cams"]["front"]
    comb = np.zeros((H, W), bool)
    rows = []
    for part, color in oracle.PART_COLORS.items():
        pts, cols = oracle.get_voxel_points_by_parts(taj["grid"], oracle.PART_COLORS, [part])
        mg = np.all(img == color, axis=-1)
        if pts.shape[0] == 0:
            rows.append((0, int(mg.sum())))
            continue
        proj = oracle.project_colored_voxels(pts, cols, np.array(p["cam_pos"]), np.array(p["target"]), p["f"], p["cx"], p["cy"], H, W)
        mp = np.all(proj == color, axis=-1)
        comb |= mp
        rows.append(((mg & mp).sum(), (mg | mp).sum()))
    cg = np.any(img != np.array(oracle.PART_COLORS["background"], np.uint8), axis=-1)
    rows.append(((cg & comb).sum(), (cg | comb).sum()))
    assert np.array_equal(np.array(rows, np.int64), g["taj_front_perpart_counts"])


def test_part_carve_on_asymmetric_grids_matches_live_reference(oracle):
    """tests/golden/partcarve_asym_golden.npz (make_golden.py partcarve_asym): the live reference's part_carve of grids
    that are not 4-way symmetric -- the inputs on which the rotated source occupancy decides a voxel's fate.  Pins the
    oracle on the branch the GPU clear pass / slab passes are tested against."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "partcarve_asym_golden.npz"))
    for tag in g["asym_cases"]:
        key = f"asym_{tag}"
        assert np.array_equal(oracle.part_carve(g[key + "_grid"], g[key + "_ext"], GROUP_JOBS), g[key + "_partcarve"]), tag


def test_portrait_monument_carving(oracle):
    """Charminar@128 (portrait mask, grid width 88) of tests/golden/real5_golden.npz: the oracle against the live
    reference's arrays and printed log (the larger cases of that file are GPU-side checks only: minutes in the oracle)."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "real5_golden.npz"))
    key = "real_Charminar_128"
    sem, ext, binm = g[key + "_sem"], g[key + "_ext"], g[key + "_bin"]
    grid = oracle.global_carve(binm, ext, 90)
    assert np.array_equal(grid, g[key + "_global"])
    assert sha(oracle.part_carve(grid, ext, GROUP_JOBS)) == str(g[key + "_partcarve_sha"])
    log = []
    out = oracle.partwise_carve(grid, ext, sem, oracle.PART_COLORS_NP, GROUP_JOBS, PART_SYMMETRY, EXTRUSION_DEPTHS, log=log)
    assert np.array_equal(out, g[key + "_partwise"])
    assert "\n".join(log) == str(g[key + "_log"]).strip("\n")
