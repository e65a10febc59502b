"""Stage hand-off formats of the reference's notebooks -- plain host I/O, no arithmetic.

    stage 1 -> 2/3   <name>_voxel_grid.npz            key `voxel_grid`, (A0,A1,A2,3) uint8     (nb1 cell 9, nb3 cell 9)
    stage 2 -> 3/4   <name>_camera_params_<tag>.json  {view: {cam_pos, target, f, cx, cy[, H, W]}}  (nb2 cell 11)

The loaders return what the consuming notebook builds from the files: notebook 2 keeps the JSON lists as float64
arrays, notebooks 3/4 convert them with dtype=float32 (nb3 cell 3 `to_numpy`), which switches the projection to
float32 -- hence the explicit `dtype` argument.
"""
from __future__ import annotations

import json
import os

import numpy as np


def save_voxel_grid(path, voxel_grid) -> str:
    """np.savez_compressed(path, voxel_grid=grid) (nb1 cell 9); accepts NumPy arrays and CUDA tensors."""
    if hasattr(voxel_grid, "detach"):
        voxel_grid = voxel_grid.detach().cpu().numpy()
    grid = np.ascontiguousarray(voxel_grid)
    if grid.dtype != np.uint8 or grid.ndim != 4 or grid.shape[3] != 3:
        raise ValueError(f"expected an (A0,A1,A2,3) uint8 grid, got {grid.dtype} {grid.shape}")
    path = os.fspath(path)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez_compressed(path, voxel_grid=grid)
    return path if path.endswith(".npz") else path + ".npz"


def load_voxel_grid(path) -> np.ndarray:
    """np.load(path)["voxel_grid"] (nb2 cell 3, nb3 cell 3)."""
    with np.load(os.fspath(path)) as z:
        return z["voxel_grid"]


def to_json_safe(obj):
    """nb2 cell 11."""
    if hasattr(obj, "detach"):
        obj = obj.detach().cpu().numpy()
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, np.generic):
        return obj.item()
    if isinstance(obj, dict):
        return {k: to_json_safe(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [to_json_safe(v) for v in obj]
    return obj


def save_camera_params(path, params) -> str:
    """json.dump(to_json_safe(params), f, indent=2) (nb2 cell 11)."""
    path = os.fspath(path)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        json.dump(to_json_safe(params), f, indent=2)
    return path


def to_numpy(obj, dtype=np.float32):
    """nb3 cell 3: lists -> arrays of `dtype` (float32 there), dicts recursively, scalars untouched."""
    if isinstance(obj, list):
        return np.array(obj, dtype=dtype)
    if isinstance(obj, dict):
        return {k: to_numpy(v, dtype) for k, v in obj.items()}
    return obj


def load_camera_params(path, dtype=np.float32):
    """Camera JSON -> nested dict with `cam_pos` / `target` as arrays of `dtype` (float32 as notebooks 3/4 do,
    float64 to continue stage 2)."""
    with open(os.fspath(path)) as f:
        return to_numpy(json.load(f), dtype)
