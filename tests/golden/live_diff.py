"""Randomised differential run of the oracle against the LIVE reference (build container only; a tool, not a test --
the test-suite stays hermetic and reads the committed fixtures).  Many more trials than the fixtures hold: every
mismatch is printed and counted, the summary goes to stdout.

    python tests/golden/live_diff.py [seed] [trials]  >  tests/golden/live_diff_r1.log
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _live_reference                                   # noqa: E402
ref = _live_reference.load()
from oracle import oracle as orc                         # noqa: E402

orc.build()
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 20240607
T = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rng = np.random.default_rng(seed)
C = ref.cfg
names = [k for k in C.PART_COLORS if k != "background"]
results = {}


def tally(name, ok):
    r = results.setdefault(name, [0, 0])
    r[0] += 1
    r[1] += 0 if ok else 1
    if not ok:
        print(f"MISMATCH {name} (trial {r[0]})")


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


# ---- camera path -----------------------------------------------------------------------------------------------
for t in range(T * 4):
    dt = np.float32 if t % 3 == 2 else np.float64
    eye = (rng.normal(0, 50, 3)).astype(dt)
    tgt = (rng.normal(0, 50, 3)).astype(dt)
    if t % 11 == 0:
        tgt = eye + np.array([0, rng.choice([-1, 1]) * rng.uniform(1, 80), 0], dtype=dt)      # |z.up| ~ 1 branch
    tally("look_at_rotation", np.array_equal(orc.look_at_rotation(eye, tgt), ref.cg.look_at_rotation(eye, tgt)))
for t in range(T):
    dt = np.float32 if t % 3 == 2 else np.float64
    n = int(rng.integers(1, 4000))
    H, W = (int(v) for v in rng.integers(8, 160, 2))
    size = rng.uniform(8, 200)
    pts = rng.integers(0, int(size) + 1, (n, 3)).astype(np.float32)
    cols = rng.integers(0, 256, (n, 3)).astype(np.uint8)
    ctr = pts.mean(0)
    eye = (ctr + rng.normal(0, 1, 3) * size * rng.choice([0.2, 1.5, 4.0])).astype(dt)
    tgt = (ctr + rng.normal(0, 1, 3) * size * 0.2).astype(dt)
    f, cx, cy = dt(rng.uniform(0.3, 4) * max(H, W)), dt(W / 2 + rng.normal(0, W / 3)), dt(H / 2 + rng.normal(0, H / 3))
    a = orc.project_colored_voxels(pts, cols, eye, tgt, f, cx, cy, H, W)
    b = ref.pu.project_colored_voxels(pts, cols, eye, tgt, f, cx, cy, H, W)
    tally(f"project_colored_voxels[{np.dtype(dt).name}]", np.array_equal(a, b))
    parts = list(rng.choice(names, size=int(rng.integers(1, 6)), replace=False))
    pc = {p: C.PART_COLORS[p] for p in parts}
    img_a = np.zeros((H, W, 3), np.uint8)
    img_b = np.zeros((H, W, 3), np.uint8)
    for im in (img_a, img_b):
        lab = rng.integers(0, len(parts) + 2, (H, W))
        for k, p in enumerate(parts):
            im[lab == k + 1] = C.PART_COLORS[p]
    da, ma = orc.compute_partwise_iou(img_a, img_b, pc)
    db, mb = ref.ce.compute_partwise_iou(img_a, img_b, pc)
    tally("compute_partwise_iou", da == db and ma == mb)
    tally("mask_parts_from_image", np.array_equal(orc.mask_parts_from_image(img_a, C.PART_COLORS, parts),
                                                  ref.mu.mask_parts_from_image(img_a, C.PART_COLORS, parts)))
    g = np.zeros((int(rng.integers(2, 14)), int(rng.integers(2, 14)), int(rng.integers(2, 14)), 3), np.uint8)
    lab = rng.integers(0, len(parts) + 2, g.shape[:3])
    for k, p in enumerate(parts):
        g[lab == k + 1] = C.PART_COLORS[p]
    pa, ca = orc.get_voxel_points_by_parts(g, C.PART_COLORS, parts)
    pb, cb = ref.vu.get_voxel_points_by_parts(g, C.PART_COLORS, parts)
    tally("get_voxel_points_by_parts", np.array_equal(pa, pb) and np.array_equal(ca, cb) and pa.dtype == pb.dtype)

# ---- carving path ----------------------------------------------------------------------------------------------
for t in range(T):
    W, H = int(rng.integers(3, 40)), int(rng.integers(2, 30))
    D = W if t % 2 == 0 else int(rng.integers(3, 40))
    vol = (rng.random((W, H, D)) < rng.uniform(0.2, 0.9)).astype(np.uint8)
    mask = (rng.random((H, W)) < 0.7).astype(np.uint8)
    interval = int(rng.choice([90, 45, 30, 60, 5, 13]))
    with quiet():
        b = ref.vc.process_voxel_grid(vol, mask, interval)
    tally(f"process_voxel_grid[{interval}]", np.array_equal(orc.process_voxel_grid(vol, mask, interval), b))
for t in range(T // 2):
    H, W = int(rng.integers(8, 48)), int(rng.integers(8, 48))
    if t % 4 == 0:
        H = W
    sem = np.empty((H, W, 3), np.uint8)
    sem[:] = C.PART_COLORS["background"]
    for _ in range(6):
        p = rng.choice(names)
        y0, x0 = int(rng.integers(0, H - 3)), int(rng.integers(0, W - 3))
        sem[y0:y0 + int(rng.integers(2, H)), x0:x0 + int(rng.integers(2, W))] = C.PART_COLORS[p]
    ext = sem.copy()
    for q in C.INTERIOR_PARTS:
        ext[np.all(sem == C.PART_COLORS_NP[q], axis=-1)] = C.PART_COLORS_NP["full_building"]
    binm = (~np.all(ext == C.PART_COLORS_NP["background"], axis=-1)).astype(np.uint8)
    ga = orc.global_carve(binm, ext, 90)
    gb = ref.vc.global_carve(binm, ext, 90)
    tally("global_carve", np.array_equal(ga, gb))
    jobs = [([n], int(rng.choice([90, 90, 90, 45]))) for n in rng.choice(names, size=4, replace=False)]
    # part_carve on a perturbed (no longer symmetric) grid
    gp = gb.copy()
    gp[rng.random(gp.shape[:3]) < 0.3] = 0
    tally("part_carve", np.array_equal(orc.part_carve(gp, ext, jobs), ref.vc.part_carve(gp, ext, jobs)))
    present = [n for n in names if np.all(gp == np.asarray(C.PART_COLORS[n], np.uint8), axis=-1).any()] or names
    col = C.PART_COLORS[str(rng.choice(present))]
    ang = int(rng.choice([5, 45, 60]))
    buf_b, log_a = io.StringIO(), []
    with contextlib.redirect_stdout(buf_b):
        lb = ref.vc.left_right_guided_carve(gp, ext, col, angle=ang)
    la = orc.left_right_guided_carve(gp, ext, col, angle=ang, log=log_a)
    tally(f"left_right_guided_carve[{ang}]", np.array_equal(la, lb))
    axis, direction = int(rng.choice([0, 2])), str(rng.choice(["+", "-"]))
    Wg, Hg, Dg = gp.shape[:3]
    m2 = rng.random((Hg, Wg)) < 0.3 if axis == 2 else rng.random((Hg, Dg)) < 0.3
    depth = int(rng.integers(1, 9))
    fill = None if t % 5 == 0 else (9, 8, 7)
    tally("extrude_from_surface", np.array_equal(orc.extrude_from_surface(gp, m2, axis, direction, depth=depth, fill_color=fill),
                                                 ref.vc.extrude_from_surface(gp, m2, axis, direction, depth=depth, fill_color=fill)))
    k, sa = int(rng.integers(1, 4)), int(rng.choice([0, 2]))
    tally("recolor_backward_components", np.array_equal(orc.recolor_backward_components(gp, col, (1, 2, 3), k=k, sort_axis=sa),
                                                        ref.vc.recolor_backward_components(gp, col, (1, 2, 3), k=k, sort_axis=sa)))

# ---- the whole notebook-1 chain on random blocky masks (notebook jobs / symmetries / extrusion depths) ---------------
sys.path.insert(0, os.path.dirname(HERE))
from helpers import EXTRUSION_DEPTHS, GROUP_JOBS, PART_SYMMETRY     # noqa: E402
for t in range(max(3, T // 6)):
    H, W = int(rng.integers(24, 56)), int(rng.integers(24, 56))
    if t % 3 == 0:
        H = W
    sem = np.empty((H, W, 3), np.uint8)
    sem[:] = C.PART_COLORS["background"]
    sem[H // 3:, W // 6: W - W // 6] = C.PART_COLORS["full_building"]
    for _ in range(8):
        p = rng.choice(names)
        y0, x0 = int(rng.integers(0, H - 3)), int(rng.integers(0, W - 3))
        sem[y0:y0 + int(rng.integers(2, H // 2)), x0:x0 + int(rng.integers(2, W // 2))] = C.PART_COLORS[p]
    ext = sem.copy()
    for q in C.INTERIOR_PARTS:
        ext[np.all(sem == C.PART_COLORS_NP[q], axis=-1)] = C.PART_COLORS_NP["full_building"]
    binm = (~np.all(ext == C.PART_COLORS_NP["background"], axis=-1)).astype(np.uint8)
    gb = ref.vc.global_carve(binm, ext, 90)
    recolor = bool(t % 2 == 0)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        pb = ref.vc.partwise_carve(gb, ext, sem, C.PART_COLORS_NP, GROUP_JOBS, PART_SYMMETRY, EXTRUSION_DEPTHS,
                                   recolor_back_minarets=recolor)
    log_a = []
    pa = orc.partwise_carve(gb, ext, sem, C.PART_COLORS_NP, GROUP_JOBS, PART_SYMMETRY, EXTRUSION_DEPTHS,
                            recolor_back_minarets=recolor, log=log_a)
    tally("partwise_carve", np.array_equal(pa, pb))
    tally("partwise_carve printed log", "\n".join(log_a) == buf.getvalue().strip("\n"))

# ---- next rows: depth-buffer visibility evaluator (eval_helpers_intra.py:134-190), hand-off points ----------------
import importlib                                          # noqa: E402
eh = importlib.import_module("utils.eval_helpers_intra")
for t in range(max(4, T // 4)):
    dt = np.float32 if t % 2 == 0 else np.float64
    A0, A1, A2 = (int(v) for v in rng.integers(4, 14, 3))
    parts = list(rng.choice(names, size=3, replace=False))
    g = np.zeros((A0, A1, A2, 3), np.uint8)
    lab = rng.integers(0, 6, (A0, A1, A2))
    for k, p in enumerate(parts):
        g[lab == k + 1] = C.PART_COLORS[p]
    H, W = (int(v) for v in rng.integers(12, 48, 2))
    ctr = np.array([A2, A1, A0]) / 2
    cam = {"cam_pos": (ctr + rng.normal(0, 1, 3) * 4 + np.array([0, 0, -3.0 * max(A0, A1, A2)])).astype(dt),
           "target": (ctr + rng.normal(0, 1, 3)).astype(dt), "f": float(rng.uniform(0.8, 2.5) * max(H, W)),
           "cx": W / 2 + float(rng.normal(0, 3)), "cy": H / 2 + float(rng.normal(0, 3))}
    zb = eh.compute_global_depth_buffer(g, cam, H, W)
    za = orc.compute_global_depth_buffer(g, cam, H, W)
    tally(f"compute_global_depth_buffer[{np.dtype(dt).name}]", np.array_equal(za, zb) and za.dtype == zb.dtype)
    pts, _ = ref.vu.get_voxel_points_by_parts(g, C.PART_COLORS, parts[:1])
    if len(pts):
        eps = float(rng.choice([1e-3, 0.75]))
        tally(f"project_part_visible[{np.dtype(dt).name}]", np.array_equal(orc.project_part_visible(pts, cam, zb, H, W, eps=eps),
                                                                            eh.project_part_visible(pts, cam, zb, H, W, eps=eps)))
    stride = int(rng.choice([1, 2, 3]))
    pa, ca = orc.voxel_grid_to_points(g, stride=stride)[:2]
    pb, cb = ref.vu.voxel_grid_to_points(g, stride=stride)[:2]
    tally("voxel_grid_to_points", np.array_equal(pa, pb) and np.array_equal(ca, cb))
    # extract_top_k_components (voxel_utils.py:22-31): speckled grid, 26-connectivity, ptp along axis 1, stable sort
    shp = tuple(int(v) for v in rng.integers(6, 22, 3))
    tg = np.zeros(shp + (3,), np.uint8)
    col = np.array(C.PART_COLORS[str(rng.choice(names))])
    tg[rng.random(shp) < rng.choice([0.08, 0.2, 0.35])] = col
    tg[rng.random(shp) < 0.1] = C.PART_COLORS["dome"]
    k = int(rng.choice([1, 2, 4]))
    tally("extract_top_k_components", np.array_equal(orc.extract_top_k_components(tg, col, k), ref.vu.extract_top_k_components(tg, col, k)))

print("\nfunction                                   trials  mismatches")
bad = 0
for k in sorted(results):
    n, m = results[k]
    bad += m
    print(f"{k:42s} {n:6d}  {m:6d}")
print(f"\nseed {seed}; numpy {np.__version__}; total mismatches: {bad}")
sys.exit(1 if bad else 0)
