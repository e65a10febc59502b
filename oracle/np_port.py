"""NumPy port of the reference's camera-scoring path, used ONLY as the timed CPU baseline
(bench.py `cpu_baseline` and `--impl reference`).  TEST/BENCH INFRASTRUCTURE, never the product.

The reference is pure NumPy; this module performs the same vectorised NumPy operations, in the same
order and dtypes, as
  look_at_rotation          utils/camera_geometry.py:3-14
  project_colored_voxels    utils/projection_utils.py:5-23
  compute_partwise_iou      utils/camera_estimation.py:770-787
  evaluate                  utils/camera_estimation.py:597-603
so that its wall time on the GPU box's host cores stands for "the reference numpy path" (the
reference itself cannot travel to the GPU box).  Its results are checked against the C oracle in
tests/test_np_port.py.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np


def look_at(eye, target):
    fwd = target - eye
    fwd = fwd / np.linalg.norm(fwd)
    up = np.array([0, 1, 0], dtype=np.float32)
    if np.allclose(np.abs(np.dot(fwd, up)), 1.0):
        up = np.array([0, 0, 1], dtype=np.float32)
    right = np.cross(up, fwd)
    right = right / np.linalg.norm(right)
    return np.stack([right, np.cross(fwd, right), fwd], axis=0)


def render(points, colours, cam_pos, target, f, cx, cy, H, W):
    cam = (points - cam_pos) @ look_at(cam_pos, target).T
    X, Y, Z = cam.T
    Z = np.where(Z < 1e-8, 1e-8, Z)
    col = np.round((X / Z) * f + cx).astype(int)
    row = np.round(-(Y / Z) * f + cy).astype(int)
    keep = (col >= 0) & (col < W) & (row >= 0) & (row < H)
    image = np.zeros((H, W, 3), dtype=np.uint8)
    image[row[keep], col[keep]] = colours[keep]
    return image


def partwise_iou(proj, gt, part_colors):
    a = proj.reshape(-1, 3)
    b = gt.reshape(-1, 3)
    counts, ious = [], []
    for colour in part_colors.values():
        pa = np.all(a == colour, axis=1)
        pb = np.all(b == colour, axis=1)
        inter = np.logical_and(pa, pb).sum()
        union = np.logical_or(pa, pb).sum()
        counts.append((int(inter), int(union)))
        ious.append(inter / union if union > 0 else 0.0)
    return counts, float(np.mean(ious))


def evaluate(points, colours, seg, part_colors, row, H, W):
    img = render(points, colours, row[0:3], row[3:6], row[6], row[7], row[8], H, W)
    return partwise_iou(img, seg, part_colors)


# ----------------------------------------------------------------------------------------------
# candidate-parallel driver (fork: the point list is shared copy-on-write)
# ----------------------------------------------------------------------------------------------
_W = {}


def _work(rows):
    out = []
    for row in rows:
        out.append(evaluate(_W["points"], _W["colours"], _W["seg"], _W["parts"], row, _W["H"], _W["W"])[1])
    return out


def timed_sweep(points, colours, seg, part_colors, cand, H, W, processes=1):
    """Score `cand` (K,9) on `processes` host processes; returns (seconds, scores)."""
    _W.update(points=points, colours=colours, seg=seg, parts=part_colors, H=H, W=W)
    cand = np.asarray(cand, dtype=np.float64)
    if processes <= 1:
        t0 = time.perf_counter()
        scores = _work(cand)
        return time.perf_counter() - t0, scores
    chunks = [c for c in np.array_split(cand, processes) if len(c)]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(chunks)) as pool:
        pool.map(_work, [c[:0] for c in chunks])            # start the workers outside the timed region
        t0 = time.perf_counter()
        parts = pool.map(_work, chunks)
        dt = time.perf_counter() - t0
    return dt, [s for p in parts for s in p]


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1
