"""Drop-in for the reference's utils/eval_helpers_intra.py: the depth-buffer visibility helpers
(:134-190), the loaders (:19-77), compute_binary_gt (:274-285) and the three table drivers of notebook 4
(run_minaret_kp_evaluation :287-424, run_minaret_iou_evaluation :427-557, run_part_minaret_binary_iou :560-748).
The matplotlib figures (`visualize=True`) are not reproduced; the tables and the DataFrames are."""
from __future__ import annotations

import json
import os
import warnings

import numpy as np
import torch

from . import _engine as eng
from . import _native as nv
from ._native import check, lib, ptr, stream_ptr
from .camera_geometry import candidate_row, working_dtype
from .projection_utils import _points_f32
from .voxel_utils import grid_to_device


def _camera_block(cam, dtype, dev):
    """Camera block in the working dtype.  look_at_rotation runs in the dtype of the camera arrays alone
    (camera_geometry.py:3-14); when the points make the projection float64 while the camera arrays are float32
    (int64 coordinates from extract_minaret_voxels_by_label), NumPy widens that float32 rotation exactly and keeps
    f, cx, cy as Python floats -- reproduced here."""
    la = working_dtype(np.asarray(cam["cam_pos"]), np.asarray(cam["target"]))
    row = candidate_row(cam["cam_pos"], cam["target"], cam["f"], cam["cx"], cam["cy"], la)
    block = eng.setup_cameras(torch.from_numpy(row[None]).to(dev))
    if la != np.dtype(dtype):
        block = block.to(torch.float64)
        block[0, 12:15] = torch.tensor([float(cam["f"]), float(cam["cx"]), float(cam["cy"])], dtype=torch.float64, device=dev)
    return block


@nv.on_device
def compute_global_depth_buffer(voxel_grid, cam, H, W, device=None, return_tensor=False):
    """eval_helpers_intra.py:134-161: float32 (H,W) buffer of the smallest camera-space Z among ALL occupied voxels
    projecting into each pixel (Z > 1e-6), +inf where nothing lands.  Arithmetic dtype follows the camera arrays
    (float32 in the reference's load_camera_json :57-77)."""
    dev = nv.require_cuda(device)
    H, W = int(H), int(W)
    g = grid_to_device(voxel_grid, dev)
    A0, A1, A2, _ = g.shape
    occ = torch.empty((A0, A1, A2), dtype=torch.uint8, device=dev)
    check(lib.p3d_crop_occupancy(ptr(g), A0, A1, A2, 0, 0, 0, A0, A1, A2, None, ptr(occ), stream_ptr()), "p3d_crop_occupancy")
    nv.launch_count += 1
    pts, _ = eng.compact_points(occ)
    dt = working_dtype(np.zeros(1, np.float32), np.asarray(cam["cam_pos"]), np.asarray(cam["target"]))
    cams = _camera_block(cam, dt, dev)
    zbuf = torch.empty((H, W), dtype=torch.float32, device=dev)
    elem = 4 if dt == np.float32 else 8
    ws_bytes = int(lib.p3d_depth_workspace_bytes(H, W, elem))
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=dev)
    fn = lib.p3d_depth_buffer_f32 if elem == 4 else lib.p3d_depth_buffer_f64
    check(fn(ptr(pts), pts.shape[0], ptr(cams), H, W, ptr(zbuf), ptr(ws), ws_bytes, stream_ptr()), "p3d_depth_buffer")
    nv.launch_count += 3
    return zbuf if return_tensor else zbuf.cpu().numpy()


@nv.on_device
def project_part_visible(pts3d, cam, zbuf, H, W, eps=1e-3, device=None, return_tensor=False):
    """eval_helpers_intra.py:168-190: boolean (H,W) mask of the pixels where one of `pts3d` lies within `eps` of the
    global depth buffer."""
    dev = nv.require_cuda(device)
    H, W = int(H), int(W)
    pts = _points_f32(pts3d, dev)
    dt = working_dtype(pts3d, np.asarray(cam["cam_pos"]), np.asarray(cam["target"]))
    cams = _camera_block(cam, dt, dev)
    zb = nv.to_device(zbuf, torch.float32, dev)
    if tuple(zb.shape) != (H, W):
        raise ValueError(f"zbuf {tuple(zb.shape)} does not match (H,W)=({H},{W})")
    mask = torch.empty((H, W), dtype=torch.uint8, device=dev)
    if dt == np.float32:
        check(lib.p3d_part_visible_f32(ptr(pts), pts.shape[0], ptr(cams), ptr(zb), float(np.float32(eps)), H, W, ptr(mask),
                                       stream_ptr()), "p3d_part_visible_f32")
    else:
        check(lib.p3d_part_visible_f64(ptr(pts), pts.shape[0], ptr(cams), ptr(zb), float(eps), H, W, ptr(mask),
                                       stream_ptr()), "p3d_part_visible_f64")
    nv.launch_count += 1
    out = mask.to(torch.bool)
    return out if return_tensor else out.cpu().numpy()


def _iou_bool(a, b):
    """eval_helpers_intra.py:275-278."""
    inter = np.logical_and(a, b).sum()
    union = np.logical_or(a, b).sum()
    return inter / union if union > 0 else np.nan


# --------------------------------------------------------------------------------------------
# loaders (:19-77)
# --------------------------------------------------------------------------------------------
def load_voxel_grid(npz_path):
    return np.load(npz_path)["voxel_grid"]


def load_mask(mask_path):
    from PIL import Image
    if not os.path.exists(mask_path):
        raise FileNotFoundError(mask_path)
    return np.array(Image.open(mask_path).convert("RGB"))


def resize_mask_to_voxel_grid(mask_img, voxel_grid):
    """:31-54: nearest-neighbour resize so that the mask's larger side equals the grid's largest extent."""
    import cv2
    H, W = mask_img.shape[:2]
    scale = max(voxel_grid.shape[:3]) / max(H, W)
    new_W, new_H = int(round(W * scale)), int(round(H * scale))
    resized = cv2.resize(mask_img, (new_W, new_H), interpolation=cv2.INTER_NEAREST)
    print(f"Mask resized: ({H},{W}) → ({new_H},{new_W}) | scale={scale:.3f}")
    return resized


def load_camera_json(path, view):
    """:56-75: float32 camera arrays, Python-float intrinsics."""
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    with open(path, "r") as f:
        data = json.load(f)
    if view not in data:
        raise KeyError(f"View '{view}' not found in {os.path.basename(path)}")
    cam = data[view]
    return {"cam_pos": np.array(cam["cam_pos"], dtype=np.float32), "target": np.array(cam["target"], dtype=np.float32),
            "f": float(cam["f"]), "cx": float(cam["cx"]), "cy": float(cam["cy"])}


def project_keypoints(voxel_kps, cam):
    """:78-82."""
    from .camera_geometry import project
    return {k: project(pt, cam["cam_pos"], cam["target"], cam["f"], cam["cx"], cam["cy"]) for k, pt in voxel_kps.items()}


@nv.on_device
def compute_binary_gt(mask_img, voxel_grid, device=None):
    """:274-285: pixels of the mask whose colour is one of the non-black colours present in the voxel grid."""
    dev = nv.require_cuda(device)
    with torch.cuda.device(dev):
        g = grid_to_device(voxel_grid, dev)
        img = nv.to_device(mask_img, torch.uint8, dev)
        present = torch.empty(int(lib.p3d_colour_presence_bytes()) // 4, dtype=torch.int32, device=dev)
        check(lib.p3d_colour_presence(ptr(g), g.numel() // 3, ptr(present), stream_ptr()), "p3d_colour_presence")
        out = torch.empty(img.shape[:2], dtype=torch.uint8, device=dev)
        check(lib.p3d_colour_lookup(ptr(img), out.numel(), ptr(present), ptr(out), stream_ptr()), "p3d_colour_lookup")
        nv.launch_count += 2
        return out.to(torch.bool).cpu().numpy()


# --------------------------------------------------------------------------------------------
# table drivers of notebook 4
# --------------------------------------------------------------------------------------------
_MINARETS = ["LM1", "RM1", "LM2", "RM2"]
_BACK_TOP_ONLY = {"Itimad": True, "Akbar": True, "Charminar": True, "Taj": False, "Bibi": False}
_MONUMENT_SHORT = {"Taj": "TM", "Bibi": "BkM", "Itimad": "IuD", "Akbar": "AT", "Charminar": "CM"}


def _no_figures(visualize):
    if visualize:
        warnings.warn("visualize=True: the matplotlib figures of the reference are not reproduced", stacklevel=3)


def _table(cells, monuments, banner):
    import pandas as pd
    from tabulate import tabulate
    df = pd.DataFrame.from_dict(cells, orient="index")
    df = df[[m for m in monuments]]
    df.columns = [_MONUMENT_SHORT[m] for m in df.columns]
    print(banner)
    print(tabulate(df, headers="keys", tablefmt="grid", showindex=True))
    return df


def _load_scene(monument, view, root_voxels, root_masks):
    grid = load_voxel_grid(os.path.join(root_voxels, f"{monument}_voxel_grid.npz"))
    mask = load_mask(os.path.join(root_masks, monument, "masks", f"{monument}_{view}_mask.png"))
    return grid, resize_mask_to_voxel_grid(mask, grid)


def run_minaret_kp_evaluation(monuments, view, root_voxels, root_masks, cam_dir, part_colors, visualize=True, device=None):
    """:287-424: minaret key-point reprojection error (px), Θinit → Θkp."""
    from . import camera_estimation as ce
    _no_figures(visualize)
    colours = [part_colors["front_minarets"], part_colors["back_minarets"]]
    cells = {m: {} for m in _MINARETS + ["Average"]}
    for monument in monuments:
        print(f"\n\U0001F3DB️ {monument}")
        grid, mask_img = _load_scene(monument, view, root_voxels, root_masks)
        cams = {"init": load_camera_json(os.path.join(cam_dir, f"{monument}_camera_params_init.json"), view),
                "rep": load_camera_json(os.path.join(cam_dir, f"{monument}_camera_params_kp.json"), view)}
        voxel_kps = ce.extract_top_bottom_voxel_points(ce.extract_minaret_voxels_by_label(grid, colours, device=device), device=device)
        image_kps = ce.extract_top_bottom_image_points(ce.extract_minaret_masks_by_label(mask_img, colours, device=device), device=device)
        err_vals = {tag: {} for tag in cams}
        for tag, cam in cams.items():
            proj_kps = project_keypoints(voxel_kps, cam)
            for m in _MINARETS:
                errs = [np.linalg.norm(np.array(image_kps[f"{m}_top"]) - np.array(proj_kps[f"{m}_top"]))]
                if not (m in ["LM2", "RM2"] and _BACK_TOP_ONLY[monument]):
                    errs.append(np.linalg.norm(np.array(image_kps[f"{m}_bottom"]) - np.array(proj_kps[f"{m}_bottom"])))
                err_vals[tag][m] = np.mean(errs)
        for m in _MINARETS:
            cells[m][monument] = f"{err_vals['init'][m]:.2f}→{err_vals['rep'][m]:.2f}"
        cells["Average"][monument] = (f"{np.mean(list(err_vals['init'].values())):.2f}"
                                      f"→{np.mean(list(err_vals['rep'].values())):.2f}")
    return _table(cells, monuments, """
=== Minaret Keypoint Reprojection Error (px) ===
Θinit → Θkp

Rules:
- LM1, RM1: top + bottom
- LM2, RM2:
    * Taj, Bibi: top + bottom
    * Akbar, Charminar, Itimad: top only
""")


def run_minaret_iou_evaluation(monuments, view, root_voxels, root_masks, cam_dir, part_colors, visualize=True, device=None):
    """:427-557: per-minaret IoU restricted to the globally visible minaret region, Θinit → Θkp → Θfinal."""
    from . import camera_estimation as ce
    _no_figures(visualize)
    colours = [part_colors["front_minarets"], part_colors["back_minarets"]]
    cells = {m: {} for m in _MINARETS + ["Average"]}
    for monument in monuments:
        print(f"\n\U0001F3DB️ {monument}")
        voxel_init, mask_img = _load_scene(monument, view, root_voxels, root_masks)
        H, W = mask_img.shape[:2]
        cams = {tag: load_camera_json(os.path.join(cam_dir, f"{monument}_camera_params_{name}.json"), view)
                for tag, name in (("init", "init"), ("rep", "kp"), ("final", "final"))}
        vox_parts = ce.extract_minaret_voxels_by_label(voxel_init, colours, device=device)
        msk_parts = ce.extract_minaret_masks_by_label(mask_img, colours, device=device)
        iou_vals = {m: {} for m in _MINARETS}
        grid_dev = grid_to_device(voxel_init, nv.require_cuda(device))
        for tag, cam in cams.items():
            zbuf = compute_global_depth_buffer(grid_dev, cam, H, W, device=device, return_tensor=True)
            pts_all = np.vstack([vox_parts[m] for m in _MINARETS])          # argwhere order, as the reference passes it
            pr_all = project_part_visible(pts_all, cam, zbuf, H, W, device=device)
            for m in _MINARETS:
                gt_m = msk_parts[m].astype(bool)
                pr_m = project_part_visible(vox_parts[m], cam, zbuf, H, W, device=device)
                iou_vals[m][tag] = _iou_bool(gt_m & pr_all, pr_m)
        for m in _MINARETS:
            cells[m][monument] = f"{iou_vals[m]['init']:.3f}→{iou_vals[m]['rep']:.3f}→{iou_vals[m]['final']:.3f}"
        cells["Average"][monument] = (f"{np.mean([iou_vals[m]['init'] for m in _MINARETS]):.3f}→"
                                      f"{np.mean([iou_vals[m]['rep'] for m in _MINARETS]):.3f}→"
                                      f"{np.mean([iou_vals[m]['final'] for m in _MINARETS]):.3f}")
    return _table(cells, monuments, """
=== Minaret IoU (INIT voxel grid)
Visualization: ALL minarets together
Table: per-minaret IoU (visible only)
Cameras: Θinit → Θkp → Θfinal
""")


def run_part_minaret_binary_iou(monuments, view, root_voxels, deformed_voxels, root_masks, cam_dir, part_colors,
                                visualize=True, device=None):
    """:560-748: visibility-aware part / minaret / whole-silhouette IoU with the final camera, init → deformed grid."""
    from .mask_utils import mask_parts_from_image
    from .voxel_utils import device_points_by_parts
    _no_figures(visualize)
    PARTS = ["dome", "chhatris", "main_door", "windows", "plinth"]
    cells = {r: {} for r in PARTS + ["minarets", "whole"]}
    dev = nv.require_cuda(device)
    for monument in monuments:
        print(f"\n\U0001F3DB️ {monument}")
        voxel_init, mask_img = _load_scene(monument, view, root_voxels, root_masks)
        voxel_def = load_voxel_grid(os.path.join(deformed_voxels, f"{monument}_deformed_voxel_grid.npz"))
        H, W = mask_img.shape[:2]
        cam = load_camera_json(os.path.join(cam_dir, f"{monument}_camera_params_final.json"), view)
        with torch.cuda.device(dev):
            g_init, g_def = grid_to_device(voxel_init, dev), grid_to_device(voxel_def, dev)
            img_dev = nv.to_device(mask_img, torch.uint8, dev)
            zbuf_init = compute_global_depth_buffer(g_init, cam, H, W, device=dev, return_tensor=True)
            zbuf_def = compute_global_depth_buffer(g_def, cam, H, W, device=dev, return_tensor=True)

            def part_gt(parts):
                return (mask_parts_from_image(img_dev, part_colors, parts, device=dev) > 0).any(dim=-1).cpu().numpy()

            for part in PARTS:
                gt = part_gt([part])
                pts_i = device_points_by_parts(g_init, part_colors, [part], dev)[0]
                pts_d = device_points_by_parts(g_def, part_colors, [part], dev)[0]
                if gt.sum() == 0 or pts_i.shape[0] == 0:
                    cells[part][monument] = "--"
                    continue
                i0 = _iou_bool(gt, project_part_visible(pts_i, cam, zbuf_init, H, W, device=dev))
                i1 = _iou_bool(gt, project_part_visible(pts_d, cam, zbuf_def, H, W, device=dev))
                cells[part][monument] = f"{i0:.3f}→{i1:.3f}"
            minaret_parts = ["front_minarets", "back_minarets"]
            pts_min = device_points_by_parts(g_init, part_colors, minaret_parts, dev)[0]   # init points for BOTH (:686-701)
            gt_min = part_gt(minaret_parts)
            i0 = _iou_bool(gt_min, project_part_visible(pts_min, cam, zbuf_init, H, W, device=dev))
            i1 = _iou_bool(gt_min, project_part_visible(pts_min, cam, zbuf_def, H, W, device=dev))
            cells["minarets"][monument] = f"{i0:.3f}→{i1:.3f}"
            gt_whole = compute_binary_gt(img_dev, g_init, device=dev)
            occ = []
            for g in (g_init, g_def):
                A0, A1, A2, _ = g.shape
                o = torch.empty((A0, A1, A2), dtype=torch.uint8, device=dev)
                check(lib.p3d_crop_occupancy(ptr(g), A0, A1, A2, 0, 0, 0, A0, A1, A2, None, ptr(o), stream_ptr()), "p3d_crop_occupancy")
                nv.launch_count += 1
                occ.append(eng.compact_points(o)[0])
            i0 = _iou_bool(gt_whole, project_part_visible(occ[0], cam, zbuf_init, H, W, device=dev))
            i1 = _iou_bool(gt_whole, project_part_visible(occ[1], cam, zbuf_def, H, W, device=dev))
            cells["whole"][monument] = f"{i0:.3f}→{i1:.3f}"
    return _table(cells, monuments, """
=== Part / Minaret / Binary IoU (init → deformed)
Camera: final (Θ*)
Visibility-aware

Binary row = true whole silhouette IoU
(not average of parts)
""")
