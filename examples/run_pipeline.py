#!/usr/bin/env python
"""Notebooks 1 -> 2 -> 3 -> 4 of the reference, end to end through the drop-in `utils` package on one B200.

    python examples/run_pipeline.py [--monument Taj] [--max-dim 256] [--out /tmp/p3d_results]

Uses the mask PNGs shipped under tests/golden/data (Taj has front + drone masks; Bibi front only).  Every call below
is the call the corresponding notebook cell makes -- only the `sys.path` line differs from the reference checkout --
except that the widget buttons are methods (`saved.aligner.run_random(...)`, `results.viewer.save_params()`).
"""
import argparse
import contextlib
import io
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "part-based-3d-reconstruction_b200"))        # instead of the reference checkout

from utils.config import INTERIOR_PARTS, PART_COLORS, PART_COLORS_NP                 # noqa: E402
from utils.mask_utils import load_and_prepare_masks, load_mask                        # noqa: E402
from utils.voxel_carving_utils import global_carve, partwise_carve                   # noqa: E402
from utils.camera_estimation import (auto_compute_initial_params_matching_bbox, extract_minaret_kps_for_view,   # noqa: E402
                                     launch_smart_aligner, optimize_camera_with_keypoints, visualize_voxel_projection_iou)
from utils.deformation_estimation import launch_deform_viewer_fixed_camera           # noqa: E402
from utils.io_utils import load_camera_params, load_voxel_grid, save_camera_params, save_voxel_grid   # noqa: E402


def timed(label, fn, silent=False):
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with (contextlib.redirect_stdout(io.StringIO()) if silent else contextlib.nullcontext()):
        out = fn()
    torch.cuda.synchronize()
    print(f"[{1e3 * (time.perf_counter() - t0):9.1f} ms] {label}")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--monument", default="Taj")
    ap.add_argument("--max-dim", type=int, default=256)
    ap.add_argument("--out", default="/tmp/p3d_results")
    ap.add_argument("--random-steps", type=int, default=500)
    a = ap.parse_args()
    data = os.path.join(ROOT, "tests", "golden", "data")
    quiet = lambda: contextlib.redirect_stdout(io.StringIO())

    # ---- notebook 1: orthographic semantic carving ------------------------------------------------------------------
    sem, sem_ext, binary = load_and_prepare_masks(data, a.monument, "front", a.max_dim, PART_COLORS_NP, INTERIOR_PARTS)
    group_jobs = [(["full_building"], 90), (["chhatris"], 90), (["plinth"], 90), (["front_minarets"], 90),
                  (["small_minarets"], 90), (["dome"], 90)]
    part_symmetry = {"dome": 5, "chhatris": 45, "front_minarets": 5, "small_minarets": 5}
    extrusion_depths = {"main_door": 20, "windows": 10}
    coloured = timed("global_carve", lambda: global_carve(binary, sem_ext, angle_interval=90))
    grid = timed("partwise_carve", lambda: partwise_carve(coloured, sem_ext, sem, PART_COLORS_NP, group_jobs, part_symmetry,
                                                          extrusion_depths), silent=True)
    print(f"             grid {grid.shape}, {int(np.count_nonzero(grid.any(-1)))} occupied voxels")
    npz = save_voxel_grid(os.path.join(a.out, "1.Orthographic_Voxel_Carving", f"{a.monument}_voxel_grid.npz"), grid)

    # ---- notebook 2: perspective camera estimation --------------------------------------------------------------------
    grid = load_voxel_grid(npz)
    front = load_mask(data, a.monument, "front", int(np.max(grid.shape)))
    parts = ["front_minarets", "back_minarets"]
    colours = [PART_COLORS[p] for p in parts]
    with quiet():
        init = auto_compute_initial_params_matching_bbox(grid, front, PART_COLORS, parts_for_alignment=parts, fov_deg=30)
    try:
        voxel_kps, image_kps = extract_minaret_kps_for_view(grid, front, colours)
        kp = timed("key-point fit (L-BFGS-B)", lambda: optimize_camera_with_keypoints(voxel_kps, image_kps, front, init),
                   silent=True)
    except ValueError as exc:                                 # monuments without four minarets: keep the bbox initialisation
        print(f"             key points skipped: {exc}")
        kp = init
    with quiet():
        saved = launch_smart_aligner(grid, front, PART_COLORS, parts_for_alignment=parts, init_params=kp)
    np.random.seed(0)
    iou = timed(f"random search, {a.random_steps} candidates", lambda: saved.aligner.run_random(a.random_steps), silent=True)
    iou = timed("coordinate descent, 5 rounds", lambda: saved.aligner.run_coord(5), silent=True)
    saved.aligner.save()
    print(f"             aligned IoU over {parts}: {iou:.4f}")
    cam_json = save_camera_params(os.path.join(a.out, "2.Perspective_Camera_Estimation", f"{a.monument}_camera_params_final.json"),
                                  {"front": dict(saved)})
    visualize_voxel_projection_iou(grid, PART_COLORS, front, dict(saved), mode="whole_on_whole")

    # ---- notebook 3: part-wise refinement with the camera fixed -------------------------------------------------------
    cam = load_camera_params(cam_json)["front"]                # float32 arrays, as notebook 3 converts them
    labels = {k: v for k, v in PART_COLORS.items() if k != "background"}
    with quiet():
        results, store = launch_deform_viewer_fixed_camera(grid, labels, front, cam, ["dome", "front_minarets"])
    for part in ("dome", "front_minarets"):
        best, best_iou = timed(f"auto-align {part} (4594 deformations)", lambda: results.viewer.run_auto_align(part), silent=True)
        if best is not None:
            with quiet():
                results.viewer.save_params()
        print(f"             {part}: {best} -> IoU {best_iou:.4f}")
    with quiet():
        deformed = results.viewer.save_deformed_grid()
    save_voxel_grid(os.path.join(a.out, "3.Part-wise_3D_Refinement", f"{a.monument}_deformed_voxel_grid.npz"), deformed)
    print(f"results under {a.out}")


if __name__ == "__main__":
    main()
