"""CPU oracle for the geometry hot path -- TEST INFRASTRUCTURE ONLY.

Restates, function by function, what the reference
(BarnitaSharma/Part-based-3D-Reconstruction, pure NumPy/SciPy) computes on the
north-star path.  Scalar numerics (projection, trilinear resample, connected
components, IoU counts) are in ``p3d_oracle.c`` with a fixed FP operation
order; the array bookkeeping around them is NumPy here.  Each function cites
the reference ``file:line`` it follows.

Who may import this module: ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py`` (``cpu_baseline`` / ``--impl reference`` legs).  The product
package never does; it raises if its CUDA library is missing.

Pinned by: ``tests/test_oracle_golden.py`` against fixtures generated from the
live reference by ``tests/golden/make_golden.py`` (run in the build container,
where ``/root/reference`` is importable).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libp3d_oracle.so")

# utils/config.py:29-40 (data, not code)
PART_COLORS = {
    "full_building": (253, 248, 96),
    "chhatris": (1, 220, 5),
    "plinth": (63, 138, 173),
    "dome": (190, 0, 255),
    "front_minarets": (0, 0, 255),
    "back_minarets": (5, 223, 223),
    "small_minarets": (255, 180, 80),
    "main_door": (180, 140, 255),
    "windows": (255, 120, 230),
    "background": (216, 224, 251),
}
PART_COLORS_NP = {k: np.array(v) for k, v in PART_COLORS.items()}
INTERIOR_PARTS = ["main_door", "windows"]


def build(force: bool = False) -> str:
    """Compile p3d_oracle.c with gcc (idempotent)."""
    src = os.path.join(_HERE, "p3d_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
        f64, f32 = ctypes.c_double, ctypes.c_float
        L.orc_look_at_f64.argtypes = [vp, vp, vp]
        L.orc_look_at_f32.argtypes = [vp, vp, vp]
        L.orc_project_f64.argtypes = [vp, vp, i64, vp, vp, f64, f64, f64, i32, i32, vp, vp]
        L.orc_project_f32.argtypes = [vp, vp, i64, vp, vp, f32, f32, f32, i32, i32, vp, vp]
        L.orc_partwise_counts.argtypes = [vp, vp, i64, vp, i32, vp, vp]
        L.orc_affine_order1_u8.argtypes = [vp, i32, i32, i32, vp, vp, vp]
        L.orc_label6.argtypes = [vp, i32, i32, i32, vp]
        L.orc_label6.restype = i32
        for fn in (L.orc_look_at_f64, L.orc_look_at_f32, L.orc_project_f64, L.orc_project_f32,
                   L.orc_partwise_counts, L.orc_affine_order1_u8):
            fn.restype = None
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------- #
# stage 2: camera scoring
# --------------------------------------------------------------------------- #
def _working_dtype(*arrs) -> np.dtype:
    """numpy promotion of `pts3d - cam_pos` etc. (projection_utils.py:7)."""
    t = np.result_type(*[np.asarray(a).dtype for a in arrs])
    return np.dtype(np.float32) if t == np.float32 else np.dtype(np.float64)


def look_at_rotation(eye, target):
    """camera_geometry.py:3-14 (default up = [0,1,0])."""
    dt = _working_dtype(eye, target)
    e = np.ascontiguousarray(eye, dtype=dt)
    t = np.ascontiguousarray(target, dtype=dt)
    R = np.empty((3, 3), dt)
    (lib().orc_look_at_f32 if dt == np.float32 else lib().orc_look_at_f64)(_p(e), _p(t), _p(R))
    return R


def project_points(pts3d, colors, cam_pos, target, f, cx, cy, H, W, want_image=True):
    """projection_utils.py:5-23.  Returns (image (H,W,3) u8 | None, winner (H,W) u32)
    where winner = 1 + index of the point that owns the pixel (0 = untouched)."""
    pts = np.ascontiguousarray(pts3d, dtype=np.float32).reshape(-1, 3)
    n = pts.shape[0]
    col = np.ascontiguousarray(colors, dtype=np.uint8).reshape(-1, 3) if colors is not None \
        else np.zeros((n, 3), np.uint8)
    dt = _working_dtype(pts, cam_pos, target)
    cp = np.ascontiguousarray(cam_pos, dtype=dt)
    tg = np.ascontiguousarray(target, dtype=dt)
    img = np.empty((H, W, 3), np.uint8) if want_image else None
    pix = np.empty((H, W), np.uint32)
    fn = lib().orc_project_f32 if dt == np.float32 else lib().orc_project_f64
    fn(_p(pts), _p(col), n, _p(cp), _p(tg), float(f), float(cx), float(cy), int(H), int(W),
       _p(img) if want_image else None, _p(pix))
    return img, pix


def project_colored_voxels(pts3d, colors, cam_pos, target, f, cx, cy, H, W):
    """projection_utils.py:5-23."""
    return project_points(pts3d, colors, cam_pos, target, f, cx, cy, H, W)[0]


def partwise_counts(proj_mask, gt_mask, part_colors: dict):
    """Integer (inter, union) pixel counts per part (camera_estimation.py:777-783)."""
    a = np.ascontiguousarray(proj_mask, dtype=np.uint8).reshape(-1, 3)
    b = np.ascontiguousarray(gt_mask, dtype=np.uint8).reshape(-1, 3)
    rgb = np.ascontiguousarray([part_colors[k] for k in part_colors], dtype=np.uint8).reshape(-1, 3)
    P = rgb.shape[0]
    inter = np.zeros(P, np.int64)
    uni = np.zeros(P, np.int64)
    lib().orc_partwise_counts(_p(a), _p(b), a.shape[0], _p(rgb), P, _p(inter), _p(uni))
    return inter, uni


def iou_from_counts(inter, uni):
    """camera_estimation.py:783-787: iou = inter/union or 0.0; score = mean over all parts."""
    ious = [float(i) / float(u) if u > 0 else 0.0 for i, u in zip(inter, uni)]
    return ious, float(np.mean(ious))


def compute_partwise_iou(proj_mask, gt_mask, part_colors: dict):
    """camera_estimation.py:770-787."""
    inter, uni = partwise_counts(proj_mask, gt_mask, part_colors)
    ious, mean = iou_from_counts(inter, uni)
    return dict(zip(part_colors.keys(), ious)), mean


def get_voxel_points_by_parts(grid, part_colors, part_names):
    """voxel_utils.py:7-21: points of the selected colours in ascending flat index,
    as float32 [x=a2, y=a1, z=a0]."""
    g = np.asarray(grid)
    sel = np.zeros(g.shape[:3], bool)
    for name in part_names:
        c = np.asarray(part_colors[name])
        sel |= (g[..., 0] == c[0]) & (g[..., 1] == c[1]) & (g[..., 2] == c[2])
    flat = np.flatnonzero(sel)
    a0, a1, a2 = np.unravel_index(flat, sel.shape)
    pts = np.stack([a2, a1, a0], axis=1).astype(np.float32)
    return pts, g.reshape(-1, 3)[flat]


def mask_parts_from_image(image, part_colors, selected_parts):
    """mask_utils.py:89-97."""
    img = np.asarray(image)
    out = np.zeros_like(img)
    for part in selected_parts:
        c = np.asarray(part_colors[part], dtype=img.dtype)
        hit = (img[..., 0] == c[0]) & (img[..., 1] == c[1]) & (img[..., 2] == c[2])
        out[hit] = c
    return out


def score_candidate(pts, cols, seg_img, selected_labels, p, H, W):
    """The `evaluate` closure, camera_estimation.py:597-603 (returns +IoU and counts)."""
    proj = project_colored_voxels(pts, cols, p["cam_pos"], p["target"], p["f"], p["cx"], p["cy"], H, W)
    inter, uni = partwise_counts(proj, seg_img, selected_labels)
    return iou_from_counts(inter, uni)[1], inter, uni


# --------------------------------------------------------------------------- #
# stage 1: carving
# --------------------------------------------------------------------------- #
def rotation_matrix_inv(angle):
    """voxel_carving_utils.py:65-69 (host NumPy/LAPACK, as in the reference)."""
    a = np.deg2rad(angle)
    c, s = np.cos(a), np.sin(a)
    return np.linalg.inv(np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]))


def affine_order1(vol_u8, M, off):
    """scipy.ndimage.affine_transform(order=1, mode='constant', cval=0) on a u8 volume."""
    v = np.ascontiguousarray(vol_u8, dtype=np.uint8)
    out = np.empty_like(v)
    Mc = np.ascontiguousarray(M, dtype=np.float64)
    oc = np.ascontiguousarray(off, dtype=np.float64)
    lib().orc_affine_order1_u8(_p(v), v.shape[0], v.shape[1], v.shape[2], _p(Mc), _p(oc), _p(out))
    return out


def mask_to_wh(mask, W, H):
    """voxel_carving_utils.py:19-28: the (H,W) test comes first, so square masks are
    always transposed."""
    if mask.shape[:2] == (H, W):
        return mask.T
    if mask.shape[:2] == (W, H):
        return mask
    raise ValueError(f"Mask shape {mask.shape} incompatible with (W,H)=({W},{H})")


def carve_with_mask(vol, mask):
    """voxel_carving_utils.py:76-87 (2-D mask branch) on a (W,H,D) occupancy volume."""
    W, H, _ = vol.shape
    m = mask_to_wh(np.asarray(mask), W, H)
    return np.where(m[:, :, None], vol, 0).astype(vol.dtype)


def process_voxel_grid(vol, mask, angle_interval=90):
    """voxel_carving_utils.py:104-126: cumulative rotate (about shape/2) + mask carve."""
    ctr = np.array(vol.shape) / 2
    out = vol
    for angle in range(0, 91, angle_interval):
        M = rotation_matrix_inv(angle)
        out = affine_order1(out, M, ctr - M @ ctr)
        out = carve_with_mask(out, mask)
    return out


def global_carve(binary_mask, semantic_mask_exterior, angle_interval=90):
    """voxel_carving_utils.py:269-298 (+ :128-136 colouring)."""
    h, w = binary_mask.shape
    carved = process_voxel_grid(np.ones((w, h, w), np.uint8), binary_mask, angle_interval)
    colour_wh = np.asarray(semantic_mask_exterior).transpose(1, 0, 2)          # (W,H,3)
    return np.where((carved == 1)[..., None], colour_wh[:, :, None, :], 0).astype(np.uint8)


def _colour_eq(arr, colour):
    c = np.asarray(colour)
    return (arr[..., 0] == c[0]) & (arr[..., 1] == c[1]) & (arr[..., 2] == c[2])


def part_carve(colored_grid, semantic_mask, group_jobs):
    """voxel_carving_utils.py:139-160."""
    final = np.zeros_like(colored_grid)
    for names, angle in group_jobs:
        in_group = np.zeros(semantic_mask.shape[:2], bool)
        for n in names:
            in_group |= _colour_eq(semantic_mask, PART_COLORS[n])
        if not in_group.any():
            continue
        m = in_group.T.astype(np.uint8)                                         # (W,H)
        sub = colored_grid * m[:, :, None, None]
        occ = (sub > 0).any(-1).astype(np.uint8)
        carved = process_voxel_grid(occ, m, angle)
        part = sub * carved[..., None]
        keep = (part > 0).any(-1)
        final[keep] = part[keep]
    return final


def label6(mask):
    """scipy.ndimage.label default structure; ids in raster order of first voxel."""
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    out = np.empty(m.shape, np.int32)
    n = lib().orc_label6(_p(m), m.shape[0], m.shape[1], m.shape[2], _p(out))
    return out, int(n)


def left_right_guided_carve(colored_grid, semantic_mask, target_color, angle=60, log=None):
    """voxel_carving_utils.py:163-210.  `log` (list) collects the lines the reference prints."""
    out = colored_grid.copy()
    in_part = _colour_eq(semantic_mask, target_color)                            # (H,W)
    if not in_part.any():
        if log is not None:
            log.append(f"[SKIP] No mask for color {target_color}")
        return out
    lab, n = label6(_colour_eq(colored_grid, target_color))
    if log is not None:
        log.append(f"[{target_color}] 3D components: {n}")
    for i in range(1, n + 1):
        member = lab == i
        idx = np.argwhere(member)
        if idx.size == 0:
            continue
        x0, y0, z0 = idx.min(0)
        x1, y1, z1 = idx.max(0) + 1
        if log is not None:
            log.append(f"  - Component {i}: bbox ({x0},{y0},{z0}) → ({x1},{y1},{z1})")
        crop2d = in_part[y0:y1, x0:x1]
        sub = colored_grid[x0:x1, y0:y1, z0:z1].copy()
        occ = (sub > 0).any(-1).astype(np.uint8)
        kept = process_voxel_grid(occ, crop2d, angle)
        if log is not None:
            log.append(f"    carved voxels: {np.count_nonzero(kept)}")
        kept_rgb = sub * kept[..., None]
        view = out[x0:x1, y0:y1, z0:z1]
        view[member[x0:x1, y0:y1, z0:z1]] = 0
        nz = (kept_rgb > 0).any(-1)
        view[nz] = kept_rgb[nz]
    return out


def extrude_from_surface(grid, mask_2d, axis, direction="+", depth=5, fill_color=None):
    """voxel_carving_utils.py:213-248 (argmax of an empty column is 0)."""
    occ = (grid > 0).any(-1)
    W, H, D = occ.shape
    fill = np.zeros((W, H, D), bool)
    sign = 1 if direction == "+" else -1
    if axis == 2:
        first = np.argmax(occ, axis=2) if sign > 0 else D - 1 - np.argmax(occ[:, :, ::-1], axis=2)
        ok2d = np.asarray(mask_2d).T.astype(bool)                                # (W,H)
        xs, ys = np.nonzero(ok2d)
        for d in range(depth):
            z = first[xs, ys] + sign * d
            k = (z >= 0) & (z < D)
            fill[xs[k], ys[k], z[k]] = True
    elif axis == 0:
        first = np.argmax(occ, axis=0) if sign > 0 else W - 1 - np.argmax(occ[::-1], axis=0)
        ok2d = np.asarray(mask_2d).astype(bool)                                  # (H, W) used as [y, z]
        if ok2d.shape != first.shape:
            raise ValueError("operands could not be broadcast together")
        ys, zs = np.nonzero(ok2d)
        for d in range(depth):
            x = first[ys, zs] + sign * d
            k = (x >= 0) & (x < W)
            fill[x[k], ys[k], zs[k]] = True
    out = grid.copy()
    out[fill] = 0 if fill_color is None else fill_color
    return out


def recolor_backward_components(voxel_grid, color, new_color, k=4, sort_axis=2):
    """voxel_carving_utils.py:252-266."""
    lab, n = label6(_colour_eq(voxel_grid, color))
    means = []
    for i in range(1, n + 1):
        idx = np.argwhere(lab == i)
        means.append((i, idx[:, sort_axis].mean()))
    keep = {i for i, _ in sorted(means, key=lambda t: t[1])[:k]}
    out = np.ascontiguousarray(voxel_grid).copy()
    for i in range(1, n + 1):
        if i not in keep:
            out[lab == i] = new_color
    return out


def partwise_carve(colored_voxel_grid, semantic_mask_exterior, semantic_mask_full, part_colors_np,
                   group_jobs, part_symmetry, extrusion_depths, recolor_back_minarets=True, log=None):
    """voxel_carving_utils.py:302-400."""
    grid = part_carve(colored_voxel_grid, semantic_mask_exterior, group_jobs)
    for part, angle in part_symmetry.items():
        grid = left_right_guided_carve(grid, semantic_mask_exterior, part_colors_np[part], angle, log)
    for part, depth in extrusion_depths.items():
        m = _colour_eq(np.asarray(semantic_mask_full), part_colors_np[part])
        for axis, direction in ((2, "+"), (2, "-"), (0, "+"), (0, "-")):
            grid = extrude_from_surface(grid, m, axis, direction, depth, part_colors_np[part])
    if recolor_back_minarets:
        oriented = np.flip(grid.transpose(2, 1, 0, 3), axis=1)
        grid = recolor_backward_components(oriented, part_colors_np["front_minarets"],
                                           part_colors_np["back_minarets"], k=2, sort_axis=0)
    return grid


# --------------------------------------------------------------------------- #
# evaluation helper next to the path: depth-buffer visibility
# --------------------------------------------------------------------------- #
def _depth_fns():
    L = lib()
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    for name, ft in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
        z = getattr(L, f"orc_depth_buffer_{name}")
        v = getattr(L, f"orc_part_visible_{name}")
        z.argtypes = [vp, i64, vp, vp, ft, ft, ft, i32, i32, vp]
        v.argtypes = [vp, i64, vp, vp, ft, ft, ft, vp, ft, i32, i32, vp]
        z.restype = v.restype = None
    return L


def _occupied_points(voxel_grid):
    g = np.asarray(voxel_grid)
    flat = np.flatnonzero((g > 0).any(-1))
    a0, a1, a2 = np.unravel_index(flat, g.shape[:3])
    return np.stack([a2, a1, a0], axis=1).astype(np.float32)


def compute_global_depth_buffer(voxel_grid, cam, H, W):
    """eval_helpers_intra.py:134-161 (min-Z buffer of every occupied voxel, float32, inf where empty)."""
    pts = np.ascontiguousarray(_occupied_points(voxel_grid))
    dt = _working_dtype(pts, cam["cam_pos"], cam["target"])
    cp, tg = np.ascontiguousarray(cam["cam_pos"], dtype=dt), np.ascontiguousarray(cam["target"], dtype=dt)
    zbuf = np.empty((H, W), np.float32)
    fn = getattr(_depth_fns(), "orc_depth_buffer_f32" if dt == np.float32 else "orc_depth_buffer_f64")
    fn(_p(pts), pts.shape[0], _p(cp), _p(tg), float(cam["f"]), float(cam["cx"]), float(cam["cy"]), int(H), int(W), _p(zbuf))
    return zbuf


def project_part_visible(pts3d, cam, zbuf, H, W, eps=1e-3):
    """eval_helpers_intra.py:168-190 (pixels where a part point lies within eps of the global depth buffer)."""
    if np.asarray(pts3d).dtype != np.float32:
        return _project_part_visible_np(np.asarray(pts3d), cam, zbuf, H, W, eps)
    pts = np.ascontiguousarray(pts3d, dtype=np.float32).reshape(-1, 3)
    dt = _working_dtype(pts, cam["cam_pos"], cam["target"])
    cp, tg = np.ascontiguousarray(cam["cam_pos"], dtype=dt), np.ascontiguousarray(cam["target"], dtype=dt)
    zb = np.ascontiguousarray(zbuf, dtype=np.float32)
    mask = np.empty((H, W), np.uint8)
    fn = getattr(_depth_fns(), "orc_part_visible_f32" if dt == np.float32 else "orc_part_visible_f64")
    fn(_p(pts), pts.shape[0], _p(cp), _p(tg), float(cam["f"]), float(cam["cx"]), float(cam["cy"]), _p(zb),
       float(np.float32(eps)) if dt == np.float32 else float(eps), int(H), int(W), _p(mask))
    return mask.astype(bool)


def _project_part_visible_np(pts3d, cam, zbuf, H, W, eps):
    """The same function for non-float32 points (the int64 argwhere coordinates notebook 4 passes, :516-526): NumPy
    promotes `pts3d - cam_pos` to float64 while the rotation stays the float32 look-at of the float32 camera arrays."""
    R = look_at_rotation(cam["cam_pos"], cam["target"])
    pc = (pts3d - cam["cam_pos"]) @ R.T
    X, Y, Z = pc.T
    ok = Z > 1e-6
    X, Y, Z = X[ok], Y[ok], Z[ok]
    ui = np.round((X / Z) * cam["f"] + cam["cx"]).astype(int)
    vi = np.round(-(Y / Z) * cam["f"] + cam["cy"]).astype(int)
    inside = (ui >= 0) & (ui < W) & (vi >= 0) & (vi < H)
    ui, vi, Z = ui[inside], vi[inside], Z[inside]
    mask = np.zeros((H, W), bool)
    hit = np.abs(Z - zbuf[vi, ui]) < eps
    mask[vi[hit], ui[hit]] = True
    return mask


def iou_bool(a, b):
    """eval_helpers_intra.py:268-271."""
    inter, union = np.logical_and(a, b).sum(), np.logical_or(a, b).sum()
    return inter / union if union > 0 else np.nan


def compute_binary_gt(mask_img, voxel_grid):
    """eval_helpers_intra.py:274-285."""
    g = np.asarray(voxel_grid).reshape(-1, 3).astype(np.uint32)
    codes = np.unique(g[:, 0] | (g[:, 1] << 8) | (g[:, 2] << 16))
    m = np.asarray(mask_img).astype(np.uint32)
    mc = m[..., 0] | (m[..., 1] << 8) | (m[..., 2] << 16)
    return np.isin(mc, codes[codes != 0])


def minaret_kp_errors(voxel_grid, mask_img, cams, minaret_colors, back_top_only):
    """Numbers behind run_minaret_kp_evaluation (eval_helpers_intra.py:346-393): {tag: {minaret: mean error in px}}."""
    vk = top_bottom_voxel_points(minaret_voxels_by_label(voxel_grid, minaret_colors))
    ik = top_bottom_image_points(minaret_masks_by_label(mask_img, minaret_colors))
    out = {}
    for tag, cam in cams.items():
        pk = {k: project_point(p, cam["cam_pos"], cam["target"], cam["f"], cam["cx"], cam["cy"]) for k, p in vk.items()}
        out[tag] = {}
        for m in ("LM1", "RM1", "LM2", "RM2"):
            errs = [np.linalg.norm(np.array(ik[f"{m}_top"]) - np.array(pk[f"{m}_top"]))]
            if not (m in ("LM2", "RM2") and back_top_only):
                errs.append(np.linalg.norm(np.array(ik[f"{m}_bottom"]) - np.array(pk[f"{m}_bottom"])))
            out[tag][m] = np.mean(errs)
    return out


def minaret_visible_ious(voxel_grid, mask_img, cams, minaret_colors):
    """Numbers behind run_minaret_iou_evaluation (eval_helpers_intra.py:500-527): {minaret: {tag: IoU}}."""
    H, W = mask_img.shape[:2]
    vp, mp = minaret_voxels_by_label(voxel_grid, minaret_colors), minaret_masks_by_label(mask_img, minaret_colors)
    names = ("LM1", "RM1", "LM2", "RM2")
    out = {m: {} for m in names}
    for tag, cam in cams.items():
        zbuf = compute_global_depth_buffer(voxel_grid, cam, H, W)
        pr_all = project_part_visible(np.vstack([vp[m] for m in names]), cam, zbuf, H, W)
        for m in names:
            out[m][tag] = iou_bool(mp[m].astype(bool) & pr_all, project_part_visible(vp[m], cam, zbuf, H, W))
    return out


def part_minaret_binary_ious(voxel_init, voxel_def, mask_img, cam, part_colors):
    """Numbers behind run_part_minaret_binary_iou (eval_helpers_intra.py:625-722): {row: (init, deformed) | None}."""
    H, W = mask_img.shape[:2]
    z_i, z_d = compute_global_depth_buffer(voxel_init, cam, H, W), compute_global_depth_buffer(voxel_def, cam, H, W)
    out = {}
    for part in ("dome", "chhatris", "main_door", "windows", "plinth"):
        gt = np.any(mask_parts_from_image(mask_img, part_colors, [part]) > 0, axis=-1)
        p_i, p_d = get_voxel_points_by_parts(voxel_init, part_colors, [part])[0], get_voxel_points_by_parts(voxel_def, part_colors, [part])[0]
        if gt.sum() == 0 or p_i.shape[0] == 0:
            out[part] = None
            continue
        out[part] = (iou_bool(gt, project_part_visible(p_i, cam, z_i, H, W)), iou_bool(gt, project_part_visible(p_d, cam, z_d, H, W)))
    mins = ["front_minarets", "back_minarets"]
    p_min = get_voxel_points_by_parts(voxel_init, part_colors, mins)[0]
    gt = np.any(mask_parts_from_image(mask_img, part_colors, mins) > 0, axis=-1)
    out["minarets"] = (iou_bool(gt, project_part_visible(p_min, cam, z_i, H, W)), iou_bool(gt, project_part_visible(p_min, cam, z_d, H, W)))
    gt = compute_binary_gt(mask_img, voxel_init)
    out["whole"] = (iou_bool(gt, project_part_visible(_occupied_points(voxel_init), cam, z_i, H, W)),
                    iou_bool(gt, project_part_visible(_occupied_points(voxel_def), cam, z_d, H, W)))
    return out


# ------------------------------------------------------------------------------------------------
# stage 3: part-wise deformation with a fixed camera (utils/deformation_estimation.py)
# ------------------------------------------------------------------------------------------------
DEFORM_OFFSETS = np.array([[0, 0, 0], [0.25, 0, 0], [-0.25, 0, 0], [0, 0.25, 0], [0, -0.25, 0],
                           [0, 0, 0.25], [0, 0, -0.25]])                    # deformation_estimation.py:87-92


def deform_coords(coords, image_shape, voxel_shape, deform):
    """`deform_coords` closure, deformation_estimation.py:70-103: seven jittered copies of the part's voxel coordinates,
    each scaled/shifted about ITS OWN mean, rounded half-even to integers; rows made unique (lexicographic x, y, z)."""
    H_img, W_img = image_shape
    D, H, W = voxel_shape
    px, py, pz = W / float(W_img), H / float(H_img), D / float(W_img)      # :76-78 (z uses the image WIDTH)
    out = []
    for off in DEFORM_OFFSETS:
        c = coords + off                                                    # :95 (float64: offsets are float64)
        centre = c.mean(axis=0, keepdims=True)                              # :72
        c = c - centre
        c[:, 0] = c[:, 0] * deform["scale_xz"] + deform["shift_xz"] * px * np.sign(c[:, 0])   # :79
        c[:, 1] = c[:, 1] * deform["scale_y"] - deform["shift_y"] * py                          # :80
        c[:, 2] = c[:, 2] * deform["scale_xz"] + deform["shift_xz"] * pz * np.sign(c[:, 2])   # :81
        out.append(np.round(c + centre).astype(int))                        # :82
    return np.unique(np.vstack(out), axis=0)                                # :100-102


def deform_valid(coords_def, voxel_shape):
    """Bounds test shared by update / save_params / save_deformed_grid (deformation_estimation.py:111-115)."""
    return ((coords_def[:, 0] >= 0) & (coords_def[:, 0] < voxel_shape[2]) &
            (coords_def[:, 1] >= 0) & (coords_def[:, 1] < voxel_shape[1]) &
            (coords_def[:, 2] >= 0) & (coords_def[:, 2] < voxel_shape[0]))


def deform_part_iou(voxel_grid, part_labels, image, cam_params, part, deform):
    """`save_params` (deformation_estimation.py:263-284): IoU of ONE part after deformation, fixed camera.
    Returns (iou, number of valid deformed voxels); with no voxel left inside the grid the projection is empty and the
    IoU is 0/|gt| = 0.0, as in the reference."""
    coords, colors = get_voxel_points_by_parts(voxel_grid, part_labels, [part])
    cd = deform_coords(coords.copy(), image.shape[:2], voxel_grid.shape[:3], deform)
    cd = cd[deform_valid(cd, voxel_grid.shape[:3])]
    cols = np.repeat(colors, repeats=max(1, int(len(cd) / len(colors)) + 1), axis=0)[:len(cd)]   # :274
    proj = project_colored_voxels(cd.astype(np.float32), cols, cam_params["cam_pos"], cam_params["target"],
                                  cam_params["f"], cam_params["cx"], cam_params["cy"], image.shape[0], image.shape[1])
    iou, _ = compute_partwise_iou(proj, image, {part: part_labels[part]})
    return float(iou[part]), len(cd)


def deformed_grid(voxel_grid, part_labels, image, saved_params):
    """`save_deformed_grid` (deformation_estimation.py:288-311): parts WITH saved parameters re-drawn at their deformed
    coordinates into an empty grid, in `part_labels` order (later parts overwrite earlier ones)."""
    out = np.zeros_like(voxel_grid, dtype=np.uint8)
    for part in part_labels:
        if part not in saved_params:
            continue
        coords, colors = get_voxel_points_by_parts(voxel_grid, part_labels, [part])
        cd = deform_coords(coords.copy(), image.shape[:2], voxel_grid.shape[:3], saved_params[part]["deform"])
        cd = cd[deform_valid(cd, voxel_grid.shape[:3])]
        if cd.size == 0:
            continue
        cols = np.repeat(colors, repeats=max(1, int(len(cd) / len(colors)) + 1), axis=0)[:len(cd)]
        out[cd[:, 2], cd[:, 1], cd[:, 0]] = cols.astype(np.uint8)
    return out


# ------------------------------------------------------------------------------------------------
# stage 2 initialisation chain (utils/camera_estimation.py:20-344): bbox init, minaret key points, key-point fit
# ------------------------------------------------------------------------------------------------
def initial_params_matching_bbox(voxel_grid, image, part_colors, parts_for_alignment, fov_deg=30):
    """auto_compute_initial_params_matching_bbox, camera_estimation.py:56-108 (dtypes as NumPy promotes them there:
    the voxel box is float32, the camera offset float64)."""
    H_img, W_img = image.shape[:2]
    pts, _ = get_voxel_points_by_parts(voxel_grid, part_colors, parts_for_alignment)
    seg = mask_parts_from_image(image, part_colors, parts_for_alignment)
    lo, hi = pts.min(axis=0), pts.max(axis=0)                                 # :65-66 float32
    centre = (lo + hi) / 2
    size = np.linalg.norm(hi - lo)                                            # :68 float32 scalar
    ys, xs = np.where(np.any(seg > 0, axis=-1))                               # :71-72
    img_lo, img_hi = np.array([xs.min(), ys.min()]), np.array([xs.max(), ys.max()])
    img_width = np.linalg.norm(img_hi - img_lo)                               # :76
    cam_pos = centre + np.array([0, 0, -size * 2.0])                          # :80
    f = H_img / (2 * np.tan(np.deg2rad(fov_deg) / 2))                         # :85
    approx = (size * f) / (size * 2.0)                                        # :88
    scale = img_width / approx                                                # :91
    return {"cam_pos": cam_pos, "target": centre, "f": f * scale, "cx": W_img / 2, "cy": H_img / 2}, scale


def label8_2d(mask):
    """skimage.measure.label(mask) for a 2-D image: 8-connectivity, ids in raster order of the first pixel
    (camera_estimation.py:263); restated with scipy.ndimage.label and a full 3x3 structure."""
    import scipy.ndimage
    return scipy.ndimage.label(np.asarray(mask) != 0, structure=np.ones((3, 3), int))


def minaret_voxels_by_label(voxel_grid, minaret_colors):
    """extract_minaret_voxels_by_label, camera_estimation.py:176-207: the four tallest (extent along axis 1)
    6-connected components of the minaret colours, named LM1/LM2/RM1/RM2 by centroid."""
    comps = []
    for colour in minaret_colors:
        lab, n = label6(_colour_eq(voxel_grid, colour))
        for cid in range(1, n + 1):
            coords = np.argwhere(lab == cid)
            if coords.size == 0:
                continue
            comps.append((coords.mean(axis=0), np.ptp(coords[:, 1]), coords))   # :187-189 (ndarray.ptp in NumPy 1.x)
    if len(comps) < 4:
        raise ValueError(f"Expected ≥4 minarets, found {len(comps)}")
    top4 = sorted(comps, key=lambda c: -c[1])[:4]                             # stable: ties keep discovery order
    cen = np.stack([c[0] for c in top4])
    order = np.argsort(cen[:, 0])
    left = sorted(order[:2], key=lambda i: cen[i, 2])
    right = sorted(order[2:], key=lambda i: cen[i, 2])
    return {"LM1": top4[left[0]][2], "LM2": top4[left[1]][2], "RM1": top4[right[0]][2], "RM2": top4[right[1]][2]}


def minaret_masks_by_label(image, minaret_colors, min_area=50):
    """extract_minaret_masks_by_label, camera_estimation.py:247-323."""
    rgb = image[:, :, :3]
    regions = []
    for ci, colour in enumerate(minaret_colors):
        lab, n = label8_2d(_colour_eq(rgb, colour))
        for cid in range(1, n + 1):
            yy, xx = np.nonzero(lab == cid)
            if len(yy) < min_area:
                continue
            regions.append({"color_idx": ci, "centroid": (yy.mean(), xx.mean()), "mask": (lab == cid).astype(np.uint8)})
    if len(regions) < 2:
        raise ValueError("Not enough minarets for camera alignment")
    regions.sort(key=lambda r: r["centroid"][1])                              # left -> right (:279)
    mid = len(regions) // 2

    def front_back(side):
        if len(side) == 1:
            return side[0], None
        side = sorted(side, key=lambda r: (r["color_idx"], r["centroid"][0]))
        return side[0], side[1]

    (lm1, lm2), (rm1, rm2) = front_back(regions[:mid]), front_back(regions[mid:])
    out = {}
    for key, reg in (("LM1", lm1), ("RM1", rm1), ("LM2", lm2), ("RM2", rm2)):  # insertion order of :317-320
        if reg:
            out[key] = reg["mask"]
    return out


def top_bottom_voxel_points(voxel_parts):
    """extract_top_bottom_voxel_points, camera_estimation.py:329-335."""
    out = {}
    for name, vox in voxel_parts.items():
        y = vox[:, 1]
        out[f"{name}_bottom"] = vox[y == y.min()].mean(axis=0)
        out[f"{name}_top"] = vox[y == y.max()].mean(axis=0)
    return out


def top_bottom_image_points(mask_parts):
    """extract_top_bottom_image_points, camera_estimation.py:338-344."""
    out = {}
    for name, m in mask_parts.items():
        yy, xx = np.nonzero(m)
        out[f"{name}_top"] = (xx[yy == yy.min()].mean(), yy.min())
        out[f"{name}_bottom"] = (xx[yy == yy.max()].mean(), yy.max())
    return out


def minaret_kps_for_view(voxel_grid, mask_img, minaret_colors):
    """extract_minaret_kps_for_view, camera_estimation.py:20-50."""
    vparts, mparts = minaret_voxels_by_label(voxel_grid, minaret_colors), minaret_masks_by_label(mask_img, minaret_colors)
    common = [k for k in vparts if k in mparts]
    if len(common) < 2:
        raise ValueError("Not enough visible minarets")
    vk, ik = top_bottom_voxel_points({k: vparts[k] for k in common}), top_bottom_image_points({k: mparts[k] for k in common})
    vsel, isel = {}, {}
    for k in vk:
        m = k.split("_")[0]
        if ("1" in m) or ("2" in m and "top" in k):
            vsel[k], isel[k] = vk[k], ik[k]
    if len(vsel) < 2:
        raise ValueError("Not enough keypoints after filtering")
    return vsel, isel


def project_point(pt3d, cam_pos, target, f, cx, cy):
    """camera_geometry.py:17-27 (un-rounded pixel coordinates of one point)."""
    R = look_at_rotation(np.asarray(cam_pos), np.asarray(target))
    X, Y, Z = (np.asarray(pt3d) - np.asarray(cam_pos)) @ R.T
    Z = max(Z, 1e-8)
    return np.array([(X / Z) * f + cx, -(Y / Z) * f + cy])


def optimize_camera_with_keypoints(voxel_kps, image_kps, image, init_params, loss_type="L2"):
    """camera_estimation.py:110-170: L-BFGS-B on the 9 camera parameters, squared (or absolute) reprojection error of
    the key points.  Returns (params, final loss)."""
    from scipy.optimize import minimize
    H, W = image.shape[:2]
    keys = list(image_kps.keys())

    def loss(x):
        total = 0
        for k in keys:
            d = project_point(voxel_kps[k], x[0:3], x[3:6], x[6], x[7], x[8]) - image_kps[k]
            total += (np.abs(d) if loss_type == "L1" else d ** 2).sum()
        return total

    x0 = [*init_params["cam_pos"], *init_params["target"], init_params["f"], init_params["cx"], init_params["cy"]]
    bounds = [(-W, 2 * W), (-H, 2 * H), (-2000, 100), (-W, 2 * W), (-H, 2 * H), (-2000, 100), (10, 2000), (0, W), (0, H)]
    res = minimize(loss, x0, bounds=bounds, method="L-BFGS-B")
    x = res.x
    return {"cam_pos": np.array(x[0:3]), "target": np.array(x[3:6]), "f": x[6], "cx": x[7], "cy": x[8]}, res.fun


# ------------------------------------------------------------------------------------------------
# hand-off helpers
# ------------------------------------------------------------------------------------------------
def voxel_grid_to_points(grid, stride=2):
    """voxel_utils.py:35-51, RGB branch: strided non-black voxels as float32 [a2,a1,a0]*stride, their colours, and the
    reference's (shape[1], shape[0], shape[2]) tuple."""
    g = np.asarray(grid)
    ds = g[::stride, ::stride, ::stride]
    a0, a1, a2 = np.nonzero(ds.any(axis=-1))
    pts = np.stack([a2, a1, a0], axis=1).astype(np.float32) * stride
    return pts, ds[a0, a1, a2], (g.shape[1], g.shape[0], g.shape[2])


# ------------------------------------------------------------------------------------------------
# x-run segments of a point list (checker of p3d_segments_*; not a reference function: the segment
# splat is an implementation detail of the CUDA sweep, the reference's point order is
# utils/voxel_utils.py:17-19)
# ------------------------------------------------------------------------------------------------
def point_segments(pts, labels, L):
    """(S,4) uint32 segment records and the number of points the form cannot represent.  Chunks = runs of <= 32 L
    list-consecutive points with equal label, equal (y, z), x increasing by exactly 1, not crossing a multiple of 32 L
    in x; a chunk of n points is dealt out column-wise to T = ceil(n / L) segments, segment r owning the points
    r, r + T, ... of the chunk.  Record = {x_first | y << 16, z | (count-1) << 16 | (T-1) << 20 | label << 26,
    list index of the first point, 0}."""
    pts = np.asarray(pts, dtype=np.float32).reshape(-1, 3)
    labels = np.asarray(labels, dtype=np.uint8).reshape(-1)
    n = len(pts)
    if n == 0:
        return np.zeros((0, 4), np.uint32), 0
    C = 32 * L
    with np.errstate(invalid="ignore"):
        ok = (labels >= 1) & (labels <= 32) & np.all((pts >= 0) & (pts <= 65535) & (pts == np.trunc(pts)), axis=1)
        xi = np.nan_to_num(pts[:, 0], nan=0.0, posinf=0.0, neginf=0.0).astype(np.int64)
    start = np.ones(n, bool)
    start[1:] = ((labels[1:] != labels[:-1]) | (pts[1:, 1] != pts[:-1, 1]) | (pts[1:, 2] != pts[:-1, 2]) |
                 (pts[1:, 0] != pts[:-1, 0] + np.float32(1)) | (xi[1:] % C == 0))
    first = np.flatnonzero(start)
    length = np.diff(np.append(first, n))
    T = (length + L - 1) // L
    chunk = np.repeat(np.arange(len(first)), T)                       # chunk of every segment
    r = np.arange(T.sum()) - np.repeat(np.cumsum(T) - T, T)           # column of every segment
    cnt = (length[chunk] - r + T[chunk] - 1) // T[chunk]
    p = np.nan_to_num(pts[first], nan=0.0, posinf=0.0, neginf=0.0).astype(np.int64)
    rec = np.zeros((len(chunk), 4), np.uint32)
    rec[:, 0] = ((p[chunk, 0] + r) & 0xffff) | ((p[chunk, 1] & 0xffff) << 16)
    rec[:, 1] = (p[chunk, 2] & 0xffff) | ((cnt - 1) << 16) | ((T[chunk] - 1) << 20) | (labels[first][chunk].astype(np.int64) << 26)
    rec[:, 2] = first[chunk] + r
    return rec, int((~ok).sum())


# ------------------------------------------------------------------------------------------------
# utils/voxel_utils.py:22-31 and :36-49 (viewer-side helpers next to the path)
# ------------------------------------------------------------------------------------------------
def extract_top_k_components(voxel_grid, color, k=4):
    """voxel_utils.py:22-31: 26-connected components of `color`; the k with the largest np.ptp along axis 1 stay (stable
    sort: ties keep the lower id), the others are blanked."""
    import scipy.ndimage
    voxel_grid = np.asarray(voxel_grid)
    mask = np.all(voxel_grid == np.asarray(color), axis=-1)
    labeled, n = scipy.ndimage.label(mask, structure=np.ones((3, 3, 3)))
    heights = []
    for i in range(1, n + 1):
        ys = np.argwhere(labeled == i)[:, 1]
        heights.append((i, int(ys.max() - ys.min())))
    top = [i for i, _ in sorted(heights, key=lambda t: -t[1])[:k]]
    out = voxel_grid.copy()
    out[mask & ~np.isin(labeled, top)] = 0
    return out


def scalar_grid_points(grid, axis="z", stride=2):
    """voxel_utils.py:36-47 for a scalar grid, up to the colormap call: (pts float32 (N,3), vals float64 (N), (H, W, D))
    with the reference's own pairing of index arrays and extents (np.where order named zs, ys, xs; 'x' divides the
    axis-2 indices by shape[0] - 1, ...)."""
    grid = np.asarray(grid)
    W, H, D = grid.shape[:3]
    mask_ds = (grid != 0)[::stride, ::stride, ::stride]
    zs, ys, xs = np.where(mask_ds)
    pts = np.stack([xs, ys, zs], axis=1).astype(np.float32) * stride
    vals = {"x": xs, "y": ys, "z": zs}[axis] / {"x": W - 1, "y": H - 1, "z": D - 1}[axis]
    return pts, vals, (H, W, D)
