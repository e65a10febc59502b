"""Drop-in for the depth-buffer visibility helpers of the reference's utils/eval_helpers_intra.py (:134-190, :275-278).
Only these are provided (the first "next" row beside the hot path); the evaluation tables and figures of
that module are out of scope."""
from __future__ import annotations

import numpy as np
import torch

from . import _engine as eng
from . import _native as nv
from ._native import check, lib, ptr, stream_ptr
from .camera_geometry import candidate_row, working_dtype
from .projection_utils import _points_f32
from .voxel_utils import grid_to_device


def _camera_block(cam, dtype, dev):
    row = candidate_row(cam["cam_pos"], cam["target"], cam["f"], cam["cx"], cam["cy"], dtype)
    return eng.setup_cameras(torch.from_numpy(row[None]).to(dev))


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
