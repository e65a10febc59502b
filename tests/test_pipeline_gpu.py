"""Notebooks 1 -> 2 -> 3 chained through the drop-in package on the GPU (real Taj masks, max_dim 96), checked against
the oracle at every hand-off: carved grid bytes, the .npz / JSON hand-off files, bbox initialisation, candidate scores of
the aligner's random search, the saved deformation IoU and the deformed grid."""
import contextlib
import io
import os

import numpy as np
import pytest

from conftest import DATA, pkg
from helpers import EXTRUSION_DEPTHS, GROUP_JOBS, PART_SYMMETRY

MAX_DIM = 96


@pytest.mark.gpu
def test_notebooks_1_to_3_chain(oracle, tmp_path):
    cfg, mu, vc = pkg("utils.config"), pkg("utils.mask_utils"), pkg("utils.voxel_carving_utils")
    ce, de, io_utils = pkg("utils.camera_estimation"), pkg("utils.deformation_estimation"), pkg("utils.io_utils")
    quiet = contextlib.redirect_stdout(io.StringIO())

    # ---- notebook 1: masks -> global_carve -> partwise_carve -> npz ------------------------------------------------
    sem, sem_ext, binary = mu.load_and_prepare_masks(DATA, "Taj", "front", MAX_DIM, cfg.PART_COLORS_NP, cfg.INTERIOR_PARTS)
    with quiet:
        coloured = vc.global_carve(binary, sem_ext, angle_interval=90)
        final = vc.partwise_carve(coloured, sem_ext, sem, cfg.PART_COLORS_NP, GROUP_JOBS, PART_SYMMETRY, EXTRUSION_DEPTHS)
        want = oracle.partwise_carve(oracle.global_carve(binary, sem_ext, 90), sem_ext, sem, oracle.PART_COLORS_NP,
                                     GROUP_JOBS, PART_SYMMETRY, EXTRUSION_DEPTHS)
    assert np.array_equal(final, want)
    path = io_utils.save_voxel_grid(tmp_path / "results_temp" / "1.Orthographic_Voxel_Carving" / "Taj_voxel_grid.npz", final)

    # ---- notebook 2: load, bbox initialisation, random search around it, JSON ---------------------------------------
    grid = io_utils.load_voxel_grid(path)
    front = mu.load_mask(DATA, "Taj", "front", int(np.max(grid.shape)))
    parts = ["front_minarets", "back_minarets"]
    with quiet:
        init = ce.auto_compute_initial_params_matching_bbox(grid, front, cfg.PART_COLORS, parts)
        init_ref, _ = oracle.initial_params_matching_bbox(grid, front, oracle.PART_COLORS, parts)
    for k in init:
        assert np.array_equal(np.asarray(init[k]), np.asarray(init_ref[k])), k
    cand = ce.random_candidates(ce.params_to_row(init), 24, np.random.default_rng(5))
    scores, counts, best = ce.score_camera_candidates(grid, front, cfg.PART_COLORS, parts, cand)
    pts, cols = oracle.get_voxel_points_by_parts(grid, oracle.PART_COLORS, parts)
    seg = oracle.mask_parts_from_image(front, oracle.PART_COLORS, parts)
    sel = {p: oracle.PART_COLORS[p] for p in parts}
    for k in (0, 7, int(best)):
        s, inter, uni = oracle.score_candidate(pts, cols, seg, sel, ce.row_to_params(cand[k]), *front.shape[:2])
        assert scores[k] == s and np.array_equal(counts[k, :, 0], inter) and np.array_equal(counts[k, :, 1], uni)
    assert best == int(np.argmax(scores))
    with quiet:
        saved = ce.launch_smart_aligner(grid, front, cfg.PART_COLORS, parts_for_alignment=parts, init_params=ce.row_to_params(cand[best]))
        saved.aligner.save()
    cam_json = io_utils.save_camera_params(tmp_path / "results_temp" / "2.Perspective_Camera_Estimation" / "Taj_camera_params_final.json",
                                           {"front": dict(saved)})

    # ---- notebook 3: float32 camera from the JSON, deform one part, save the deformed grid ---------------------------
    cam = io_utils.load_camera_params(cam_json)["front"]
    assert cam["cam_pos"].dtype == np.float32
    labels = {k: v for k, v in cfg.PART_COLORS.items() if k != "background"}
    deform = {"scale_y": 1.05, "shift_y": 2.0, "scale_xz": 0.95, "shift_xz": -1.0}
    with quiet:
        results, store = de.launch_deform_viewer_fixed_camera(grid, labels, front, cam, ["dome", "front_minarets"])
        results.viewer.set_sliders(part="dome", **deform)
        iou = results.viewer.save_params()
        deformed = results.viewer.save_deformed_grid()
    want_iou, _ = oracle.deform_part_iou(grid, labels, front, cam, "dome", deform)
    assert iou == want_iou == results["dome"]["iou"]
    assert np.array_equal(deformed, oracle.deformed_grid(grid, labels, front, {"dome": results["dome"]}))
    p3 = io_utils.save_voxel_grid(tmp_path / "results_temp" / "3.Part-wise_3D_Refinement" / "Taj_deformed_voxel_grid.npz", store["grid"])
    assert np.array_equal(io_utils.load_voxel_grid(p3), deformed)
