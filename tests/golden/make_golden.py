"""Generate the committed golden fixtures from the LIVE reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the unmodified reference (tests/golden/_live_reference.py stubs the plotting
modules), runs the hot-path functions on real and synthetic inputs, and writes
  tests/golden/data/...            input assets copied from the reference's data/ and results/
                                   (mask PNGs, the stored Taj voxel grid, Taj camera JSONs)
  tests/golden/camera_golden.npz   look_at / projection / IoU outputs
  tests/golden/carve_golden.npz    affine / label / global_carve / partwise_carve outputs
Nothing under tests/ reads /root/reference at test time.
"""
import contextlib
import hashlib
import io
import json
import os
import shutil
import sys

import numpy as np
import scipy.ndimage

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _live_reference  # noqa: E402

ref = _live_reference.load()
C = ref.cfg
DATA = os.path.join(HERE, "data")
SEED = 20240607

GROUP_JOBS = [(["full_building"], 90), (["chhatris"], 90), (["plinth"], 90), (["front_minarets"], 90),
              (["small_minarets"], 90), (["dome"], 90)]                                   # nb1 cell 7
PART_SYMMETRY = {"dome": 5, "chhatris": 45, "front_minarets": 5, "small_minarets": 5}
EXTRUSION_DEPTHS = {"main_door": 20, "windows": 10}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def copy_assets():
    for name, views in (("Bibi", ["front"]), ("Taj", ["front", "drone"]), ("Akbar", ["front"])):
        d = os.path.join(DATA, name, "masks")
        os.makedirs(d, exist_ok=True)
        for v in views:
            fn = f"{name}_{v}_mask.png"
            shutil.copyfile(os.path.join(ref.root, "data", name, "masks", fn), os.path.join(d, fn))
    r1 = os.path.join(DATA, "results", "1.Orthographic_Voxel_Carving")
    r2 = os.path.join(DATA, "results", "2.Perspective_Camera_Estimation")
    os.makedirs(r1, exist_ok=True)
    os.makedirs(r2, exist_ok=True)
    shutil.copyfile(os.path.join(ref.root, "results/1.Orthographic_Voxel_Carving/Taj_voxel_grid.npz"),
                    os.path.join(r1, "Taj_voxel_grid.npz"))
    for tag in ("init", "kp", "final"):
        fn = f"Taj_camera_params_{tag}.json"
        shutil.copyfile(os.path.join(ref.root, "results/2.Perspective_Camera_Estimation", fn), os.path.join(r2, fn))
    for root, _, files in os.walk(DATA):
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)


def perturb(base, K, rng):
    """Index 0 = base, then base + U(-1,1)*steps in the reference's draw order (camera_estimation.py:619-625)."""
    steps = np.array([50, 50, 100, 50, 50, 100, 50, 20, 20], float)
    out = np.empty((K, 9))
    out[0] = base
    out[1:] = base + rng.uniform(-1, 1, (K - 1, 9)) * steps
    return out


def ref_score(pts, cols, seg, sel, row, H, W, dtype=np.float64):
    cp, tg = row[0:3].astype(dtype), row[3:6].astype(dtype)
    f, cx, cy = (dtype(row[6]), dtype(row[7]), dtype(row[8])) if dtype is np.float32 else (row[6], row[7], row[8])
    proj = ref.pu.project_colored_voxels(pts, cols, cp, tg, f, cx, cy, H, W)
    per, mean = ref.ce.compute_partwise_iou(proj, seg, sel)
    counts = []
    for c in sel.values():
        a = np.all(proj.reshape(-1, 3) == c, axis=1)
        b = np.all(seg.reshape(-1, 3) == c, axis=1)
        counts.append(((a & b).sum(), (a | b).sum()))
    return proj, np.array(counts, np.int64), float(mean)


def camera_golden():
    out = {}
    rng = np.random.default_rng(SEED)
    # --- look_at vectors -----------------------------------------------------------------------
    eyes = rng.uniform(-800, 800, (256, 3))
    tgts = rng.uniform(-300, 600, (256, 3))
    tgts[::16, 0] = eyes[::16, 0]
    tgts[::16, 2] = eyes[::16, 2]                      # straight up/down: the allclose branch
    out["lookat_eye"], out["lookat_target"] = eyes, tgts
    out["lookat_R64"] = np.stack([ref.cg.look_at_rotation(e.copy(), t.copy()) for e, t in zip(eyes, tgts)])
    out["lookat_R32"] = np.stack([ref.cg.look_at_rotation(e.astype(np.float32), t.astype(np.float32))
                                  for e, t in zip(eyes, tgts)])
    assert out["lookat_R32"].dtype == np.float32

    # --- Taj (config 2): stored grid, front + drone masks, stored cameras ---------------------------
    grid = np.load(os.path.join(ref.root, "results/1.Orthographic_Voxel_Carving/Taj_voxel_grid.npz"))["voxel_grid"]
    cams = json.load(open(os.path.join(ref.root, "results/2.Perspective_Camera_Estimation/Taj_camera_params_final.json")))
    masks = {"front": ref.mu.load_mask(os.path.join(ref.root, "data"), "Taj", "front", int(np.max(grid.shape))),
             "drone": ref.mu.load_mask(os.path.join(ref.root, "data"), "Taj", "drone")}
    minarets = ["front_minarets", "back_minarets"]
    allparts = [p for p in C.PART_COLORS if p != "background"]
    for view in ("front", "drone"):
        img = masks[view]
        H, W = img.shape[:2]
        base = np.array([*cams[view]["cam_pos"], *cams[view]["target"], cams[view]["f"], cams[view]["cx"], cams[view]["cy"]])
        for tag, parts, K in (("min", minarets, 24), ("all", allparts, 4)):
            cand = perturb(base, K, rng)
            seg = ref.mu.mask_parts_from_image(img, C.PART_COLORS, parts)
            sel = {p: C.PART_COLORS[p] for p in parts}
            pts, cols = ref.vu.get_voxel_points_by_parts(grid, C.PART_COLORS, parts)
            counts, scores, hashes = [], [], []
            for k in range(K):
                proj, c, s = ref_score(pts, cols, seg, sel, cand[k], H, W)
                counts.append(c); scores.append(s); hashes.append(sha(proj))
                if k == 0:
                    out[f"taj_{view}_{tag}_image0"] = proj
            key = f"taj_{view}_{tag}"
            out[key + "_cand"] = cand
            out[key + "_counts"] = np.array(counts)
            out[key + "_scores"] = np.array(scores)
            out[key + "_sha"] = np.array(hashes)
            out[key + "_npts"] = np.array(pts.shape[0])
            print(key, "pts", pts.shape[0], "base score", scores[0])
        # float32 path (notebooks 3/4 pass float32 camera arrays)
        parts = minarets
        cand = perturb(base, 6, rng)
        seg = ref.mu.mask_parts_from_image(img, C.PART_COLORS, parts)
        sel = {p: C.PART_COLORS[p] for p in parts}
        pts, cols = ref.vu.get_voxel_points_by_parts(grid, C.PART_COLORS, parts)
        counts, hashes = [], []
        for k in range(6):
            proj, c, s = ref_score(pts, cols, seg, sel, cand[k], H, W, dtype=np.float32)
            counts.append(c); hashes.append(sha(proj))
        out[f"taj_{view}_f32_cand"] = cand.astype(np.float32)
        out[f"taj_{view}_f32_counts"] = np.array(counts)
        out[f"taj_{view}_f32_sha"] = np.array(hashes)

    # --- per-part scoring core of visualize_voxel_projection_iou (camera_estimation.py:381-403, 433-447)
    img = masks["front"]
    H, W = img.shape[:2]
    p = cams["front"]
    bg = np.array(C.PART_COLORS["background"], np.uint8)
    comb_prj = np.zeros((H, W), bool)
    rows = []
    for part, color in C.PART_COLORS.items():
        pts, col = ref.vu.get_voxel_points_by_parts(grid, C.PART_COLORS, [part])
        if pts.shape[0] == 0:
            rows.append((0, int(np.all(img == color, axis=-1).sum())))
            continue
        proj = ref.pu.project_colored_voxels(pts, col, np.array(p["cam_pos"]), np.array(p["target"]), p["f"], p["cx"], p["cy"], H, W)
        mg, mp = np.all(img == color, axis=-1), np.all(proj == color, axis=-1)
        comb_prj |= mp
        rows.append(((mg & mp).sum(), (mg | mp).sum()))
    comb_gt = np.any(img != bg, axis=-1)
    rows.append(((comb_gt & comb_prj).sum(), (comb_gt | comb_prj).sum()))
    out["taj_front_perpart_counts"] = np.array(rows, np.int64)
    out["taj_front_perpart_parts"] = np.array(list(C.PART_COLORS.keys()))

    # --- adversarial synthetic: camera inside the cloud (points behind the camera), tiny image, empty part
    g = np.zeros((12, 10, 14, 3), np.uint8)
    r2 = np.random.default_rng(7)
    lab = r2.integers(0, 4, g.shape[:3])
    g[lab == 1] = C.PART_COLORS["dome"]
    g[lab == 2] = C.PART_COLORS["plinth"]
    parts = ["dome", "plinth", "windows"]                 # windows: absent from grid and image
    img = np.zeros((20, 24, 3), np.uint8)
    img[4:15, 5:20] = C.PART_COLORS["dome"]
    img[12:18, 2:22] = C.PART_COLORS["plinth"]
    seg = ref.mu.mask_parts_from_image(img, C.PART_COLORS, parts)
    sel = {q: C.PART_COLORS[q] for q in parts}
    pts, cols = ref.vu.get_voxel_points_by_parts(g, C.PART_COLORS, parts)
    cand = np.array([[7.0, 5.0, 6.0, 7.0, 5.0, 30.0, 20.0, 12.0, 10.0],      # inside the cloud
                     [7.0, 5.0, -20.0, 7.0, 5.0, 6.0, 25.0, 12.0, 10.0],     # in front
                     [7.0, 40.0, 6.0, 7.0, 5.0, 6.0, 25.0, 12.0, 10.0],      # straight down (allclose branch)
                     [7.0, 5.0, 6.0, 7.0, 5.0, 6.0, 25.0, 12.0, 10.0],       # eye == target -> NaN
                     [6.5, 4.5, -15.0, 6.5, 4.5, 6.0, 12.0, 11.5, 9.5]])     # exact .5 pixel ties
    counts, hashes, scores = [], [], []
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k in range(len(cand)):
            proj, c, s = ref_score(pts, cols, seg, sel, cand[k], 20, 24)
            counts.append(c); hashes.append(sha(proj)); scores.append(s)
    out["adv_grid"], out["adv_image"], out["adv_cand"] = g, img, cand
    out["adv_counts"], out["adv_sha"], out["adv_scores"] = np.array(counts), np.array(hashes), np.array(scores)
    np.savez_compressed(os.path.join(HERE, "camera_golden.npz"), **out)


def blocky_semantic(rng, H, W, last_col=False):
    """Random blocky semantic mask in the PART_COLORS palette (background elsewhere)."""
    names = ["full_building", "chhatris", "plinth", "dome", "front_minarets", "small_minarets", "main_door", "windows"]
    sem = np.empty((H, W, 3), np.uint8)
    sem[:] = C.PART_COLORS["background"]
    for _ in range(14):
        n = names[rng.integers(len(names))]
        y0, x0 = rng.integers(0, H - 4), rng.integers(0, W - 4)
        y1, x1 = y0 + rng.integers(3, max(4, H // 2)), x0 + rng.integers(3, max(4, W // 2))
        sem[y0:y1, x0:x1] = C.PART_COLORS[n]
    if last_col:
        sem[H // 4: H // 2, W - 3:] = C.PART_COLORS["plinth"]
    return sem


def carve_golden():
    out = {}
    rng = np.random.default_rng(SEED)
    # --- scipy affine_transform / label unit vectors ------------------------------------------------
    k = 0
    for shape in [(43, 47, 44), (30, 58, 30), (16, 5, 16), (15, 4, 15), (31, 3, 31), (20, 6, 33)]:
        for ang in (0, 5, 45, 60, 90):
            vol = (rng.random(shape) < 0.55).astype(np.uint8)
            M = ref.vc._rotation_matrix_inv(ang)
            ctr = np.array(shape) / 2
            res = scipy.ndimage.affine_transform(vol, M, offset=ctr - M @ ctr, order=1, mode="constant", cval=0)
            out[f"aff{k}_vol"], out[f"aff{k}_angle"], out[f"aff{k}_out"] = np.packbits(vol), np.array(ang), np.packbits(res)
            out[f"aff{k}_shape"] = np.array(shape)
            assert res.max() <= 1
            k += 1
    out["aff_n"] = np.array(k)
    for i, (shape, p) in enumerate([((20, 30, 25), 0.3), ((24, 24, 24), 0.55), ((9, 40, 17), 0.7)]):
        m = rng.random(shape) < p
        lab, n = scipy.ndimage.label(m)
        out[f"lab{i}_mask"], out[f"lab{i}_shape"] = np.packbits(m), np.array(shape)
        out[f"lab{i}_out"], out[f"lab{i}_n"] = lab.astype(np.int32), np.array(n)
    out["lab_n"] = np.array(3)

    # --- real masks ---------------------------------------------------------------------------------
    cases = [("Bibi", 64), ("Taj", 96), ("Akbar", 128), ("Bibi", 256)]
    for name, md in cases:
        sem, ext, binm = ref.mu.load_and_prepare_masks(os.path.join(ref.root, "data"), name, "front", md,
                                                       C.PART_COLORS_NP, C.INTERIOR_PARTS)
        g = ref.vc.global_carve(binm, ext, 90)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            p = ref.vc.partwise_carve(g, ext, sem, C.PART_COLORS_NP, GROUP_JOBS, PART_SYMMETRY, EXTRUSION_DEPTHS)
        key = f"real_{name}_{md}"
        out[key + "_sem"], out[key + "_ext"], out[key + "_bin"] = sem, ext, binm
        out[key + "_global_sha"], out[key + "_partwise_sha"] = np.array(sha(g)), np.array(sha(p))
        out[key + "_global_occ"] = np.array(np.count_nonzero(g.any(-1)))
        out[key + "_partwise_occ"] = np.array(np.count_nonzero(p.any(-1)))
        out[key + "_log"] = np.array(buf.getvalue())
        if md <= 128:
            out[key + "_global"], out[key + "_partwise"] = g, p
        print(key, sem.shape, out[key + "_partwise_sha"], out[key + "_partwise_occ"])
    out["real_cases"] = np.array([f"{n}_{m}" for n, m in cases])

    # --- synthetic quirk cases: square image (_mask_to_wh), last-column foreground, W in the odd-offset set --
    syn = [("sq64", 64, 64, False), ("rect40x64", 40, 64, False), ("sq63", 63, 63, True), ("rect50x31", 50, 31, True),
           ("rect33x48", 33, 48, True)]
    for tag, H, W, last in syn:
        sem = blocky_semantic(rng, H, W, last)
        ext = sem.copy()
        for q in C.INTERIOR_PARTS:
            ext[np.all(sem == C.PART_COLORS_NP[q], axis=-1)] = C.PART_COLORS_NP["full_building"]
        binm = (~np.all(ext == C.PART_COLORS_NP["background"], axis=-1)).astype(np.uint8)
        g = ref.vc.global_carve(binm, ext, 90)
        pc = ref.vc.part_carve(g, ext, GROUP_JOBS)
        with contextlib.redirect_stdout(io.StringIO()):
            p = ref.vc.partwise_carve(g, ext, sem, C.PART_COLORS_NP, GROUP_JOBS, PART_SYMMETRY, EXTRUSION_DEPTHS)
            p2 = ref.vc.partwise_carve(g, ext, sem, C.PART_COLORS_NP, GROUP_JOBS, PART_SYMMETRY, EXTRUSION_DEPTHS,
                                       recolor_back_minarets=False)
        key = "syn_" + tag
        out[key + "_sem"], out[key + "_ext"], out[key + "_bin"] = sem, ext, binm
        out[key + "_global"], out[key + "_partcarve"], out[key + "_partwise"] = g, pc, p
        out[key + "_partwise_norecolor_sha"] = np.array(sha(p2))
        print(key, g.shape, np.count_nonzero(g.any(-1)), np.count_nonzero(p.any(-1)))
    out["syn_cases"] = np.array([s[0] for s in syn])
    np.savez_compressed(os.path.join(HERE, "carve_golden.npz"), **out)




def aligner_golden():
    """Drive the live launch_smart_aligner (camera_estimation.py:479-768) through fake widgets and record where the
    three optimisers end up."""
    import _fake_widgets as fw
    ce = ref.ce
    ce.widgets = fw
    ce.display = lambda *a, **k: None
    ce.clear_output = lambda *a, **k: None
    grid = np.load(os.path.join(ref.root, "results/1.Orthographic_Voxel_Carving/Taj_voxel_grid.npz"))["voxel_grid"]
    grid = np.ascontiguousarray(grid[::2, ::2, ::2])               # 256x139x256: keeps the run short
    front = ref.mu.load_mask(os.path.join(ref.root, "data"), "Taj", "front", int(np.max(grid.shape)))
    kp = json.load(open(os.path.join(ref.root, "results/2.Perspective_Camera_Estimation/Taj_camera_params_kp.json")))["front"]
    init = {"cam_pos": np.array(kp["cam_pos"]) / 2, "target": np.array(kp["target"]) / 2, "f": kp["f"] / 2,
            "cx": kp["cx"] / 2, "cy": kp["cy"] / 2}
    parts = ["front_minarets", "back_minarets"]
    out = {"grid": grid, "image": front, "init": np.array([*init["cam_pos"], *init["target"], init["f"], init["cx"], init["cy"]])}

    def snapshot(sliders):
        return np.array([sliders[k].value for k in ["cam_x", "cam_y", "cam_z", "target_x", "target_y", "target_z", "f", "cx", "cy"]])

    for lock in (False, True):
        fw.Button.instances.clear()
        captured = {}
        orig_vbox = fw.VBox

        def grab(children):
            captured["layout"] = children
            return orig_vbox(children)

        fw.VBox = grab
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            saved = ce.launch_smart_aligner(grid, front, C.PART_COLORS, parts_for_alignment=parts,
                                            init_params={k: (v.copy() if hasattr(v, "copy") else v) for k, v in init.items()},
                                            lock_xy_equal=lock)
        fw.VBox = orig_vbox
        btn = {b.description: b for b in fw.Button.instances}
        rows = captured["layout"]
        sliders = {}
        for row in rows[:4]:
            for w in row.children:
                sliders[w.description] = w
        steps = {w.description: w for w in rows[5].children}
        tag = "lock" if lock else "free"
        np.random.seed(1234)
        steps["Random Steps"].value = 12
        with contextlib.redirect_stdout(buf):
            btn["Random Search"].click()
        out[f"{tag}_after_random"] = snapshot(sliders)
        steps["Coord Steps"].value = 4
        with contextlib.redirect_stdout(buf):
            btn["Coordinate Descent"].click()
        out[f"{tag}_after_coord"] = snapshot(sliders)
        steps["Powell MaxIter"].value = 2
        with contextlib.redirect_stdout(buf):
            btn["Powell"].click()
        out[f"{tag}_after_powell"] = snapshot(sliders)
        with contextlib.redirect_stdout(buf):
            btn["Save"].click()
        out[f"{tag}_saved"] = np.array([*saved["cam_pos"], *saved["target"], saved["f"], saved["cx"], saved["cy"]])
        out[f"{tag}_log"] = np.array("\n".join(l for l in buf.getvalue().split("\n") if "Done" in l))
        print(tag, out[f"{tag}_log"])
    np.savez_compressed(os.path.join(HERE, "aligner_golden.npz"), **out)


def deform_golden():
    """Drive the live launch_deform_viewer_fixed_camera (deformation_estimation.py:15-356) through fake widgets: for a
    few (part, slider) settings record the IoU its "Save Params" button stores and the grid its "Save Deformed Grid"
    button builds -- once with float64 camera arrays and once with float32 ones (notebook 3 converts the camera JSON
    with dtype=float32, which switches the projection to float32).  Same half-resolution Taj grid / front mask /
    camera as aligner_golden.npz."""
    import _fake_widgets as fw
    import utils.deformation_estimation as de
    de.widgets = fw
    de.display = lambda *a, **k: None
    de.clear_output = lambda *a, **k: None
    de.visualize_voxel_projection_iou = lambda *a, **k: None      # figures only (deformation_estimation.py:137)
    ag = np.load(os.path.join(HERE, "aligner_golden.npz"))
    grid, front, row = ag["grid"], ag["image"], ag["free_saved"]
    part_labels = {k: v for k, v in C.PART_COLORS.items() if k != "background"}
    out = {"cam": row}
    buf = io.StringIO()

    def run(tag, cam, parts, settings):
        fw.Button.instances.clear()
        captured = {}
        orig_vbox = fw.VBox

        def grab(children):
            captured["layout"] = children
            return orig_vbox(children)

        fw.VBox = grab
        with contextlib.redirect_stdout(buf):
            saved, storage = de.launch_deform_viewer_fixed_camera(grid, part_labels, front, cam, parts)
        fw.VBox = orig_vbox
        btn = {b.description: b for b in fw.Button.instances}
        rows = captured["layout"]
        drop = rows[0]
        sl = {w.description: w for r in rows[1:3] for w in r.children}
        names, deforms, ious = [], [], []
        with contextlib.redirect_stdout(buf):
            for part in parts:
                drop.value = part
                for (sy, dy, sxz, dxz) in settings[part]:
                    sl["scale_y"].value, sl["shift_y"].value = sy, dy
                    sl["scale_xz"].value, sl["shift_xz"].value = sxz, dxz
                    btn["Save Params"].click()
                    names.append(part); deforms.append([sy, dy, sxz, dxz]); ious.append(saved[part]["iou"])
            for part in parts:                                  # keep the first setting of every part for the grid
                sy, dy, sxz, dxz = settings[part][0]
                saved[part] = {"deform": {"scale_y": sy, "shift_y": dy, "scale_xz": sxz, "shift_xz": dxz}, "iou": 0.0}
            btn["Save Deformed Grid"].click()
        g = storage["grid"]
        nz = np.flatnonzero(g.any(-1))
        out.update({f"{tag}_parts": np.array(names), f"{tag}_deforms": np.array(deforms), f"{tag}_ious": np.array(ious),
                    f"{tag}_grid_sha": np.array(sha(g)), f"{tag}_grid_nz": nz.astype(np.int64),
                    f"{tag}_grid_rgb": g.reshape(-1, 3)[nz]})
        print(tag, list(zip(names, ious)), "grid occupied", len(nz))
        return btn

    cam64 = {"cam_pos": row[0:3].copy(), "target": row[3:6].copy(), "f": float(row[6]), "cx": float(row[7]), "cy": float(row[8])}
    btn = run("f64", cam64, ["front_minarets", "dome", "plinth", "chhatris"],
              {"front_minarets": [(1.0, 0.0, 1.0, 0.0), (1.07, -6.0, 0.93, 4.0), (0.81, 11.0, 1.26, -9.0)],
               "dome": [(1.13, 5.0, 1.1, 3.0), (0.9, -12.0, 0.77, 0.0)],
               "plinth": [(1.5, 0.0, 0.6, -21.0)],
               "chhatris": [(0.5, 100.0, 2.0, 100.0), (1.0, -100.0, 1.0, 0.0)]})    # last ones leave few / no voxels inside
    cam32 = {"cam_pos": row[0:3].astype(np.float32), "target": row[3:6].astype(np.float32), "f": float(row[6]),
             "cx": float(row[7]), "cy": float(row[8])}
    run("f32", cam32, ["back_minarets", "dome"],
        {"back_minarets": [(1.0, 0.0, 1.0, 0.0), (0.95, 3.0, 1.11, -5.0)], "dome": [(1.02, -2.0, 0.97, 1.0)]})
    # deform_coords itself (closure): recover it from the callback's closure cells
    fn = btn["Save Params"]._cb
    cells = dict(zip(fn.__code__.co_freevars, fn.__closure__))
    deform_coords = cells["deform_coords"].cell_contents
    pts, _ = ref.vu.get_voxel_points_by_parts(grid, part_labels, ["back_minarets"])
    d0 = {"scale_y": 1.21, "shift_y": 7.0, "scale_xz": 0.88, "shift_xz": -13.0}
    cd = deform_coords(pts.copy(), front.shape[:2], grid.shape[:3], d0)
    out.update(coords_part=np.array("back_minarets"), coords_deform=np.array([1.21, 7.0, 0.88, -13.0]), coords_def=cd.astype(np.int32))
    np.savez_compressed(os.path.join(HERE, "deform_golden.npz"), **out)
    print("coords", cd.shape)


def init_golden():
    """Camera initialisation chain of the live reference (camera_estimation.py:20-344) on the half-resolution Taj grid
    and front mask: bbox init, minaret extraction (3-D and 2-D), top/bottom key points, key-point L-BFGS-B fit.
    Two shims, needed only because of this container's library versions (the reference's own code is unmodified):
      * ndarray.ptp() was removed in NumPy 2 (camera_estimation.py:189): np.argwhere is wrapped to return an ndarray
        subclass that still has it;
      * scikit-image is not installed: `label2d` / `regionprops` (camera_estimation.py:8, 246, 263-266) are replaced by
        scipy.ndimage.label with a full 3x3 structure (skimage's default 8-connectivity, same raster-order ids) and a
        minimal regionprops (area, centroid = mean row/col, label)."""
    import types
    ce = ref.ce

    class _Arr(np.ndarray):
        def ptp(self, *a, **k):
            return np.ptp(np.asarray(self), *a, **k)

    npx = types.ModuleType("numpy_with_ptp")
    npx.__dict__.update(np.__dict__)
    npx.argwhere = lambda m: np.argwhere(m).view(_Arr)
    ce.np = npx

    def label2d(mask):
        return scipy.ndimage.label(np.asarray(mask) != 0, structure=np.ones((3, 3), int))[0]

    class _Region:
        def __init__(self, lab, cid):
            yy, xx = np.nonzero(lab == cid)
            self.area, self.centroid, self.label = len(yy), (yy.mean(), xx.mean()), cid

    ce.label2d = label2d
    ce.regionprops = lambda lab: [_Region(lab, c) for c in range(1, int(lab.max()) + 1)]

    ag = np.load(os.path.join(HERE, "aligner_golden.npz"))
    grid, front = ag["grid"], ag["image"]
    parts = ["front_minarets", "back_minarets"]
    colours = [C.PART_COLORS[p] for p in parts]
    out = {}
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        init = ce.auto_compute_initial_params_matching_bbox(grid, front, C.PART_COLORS, parts)
    out["init_row"] = np.array([*init["cam_pos"], *init["target"], init["f"], init["cx"], init["cy"]], dtype=np.float64)
    out["init_dtypes"] = np.array([str(np.asarray(init[k]).dtype) for k in ("cam_pos", "target", "f", "cx", "cy")])
    out["init_log"] = np.array(buf.getvalue())
    vparts = ce.extract_minaret_voxels_by_label(grid, colours)
    mparts = ce.extract_minaret_masks_by_label(front, colours)
    for k, v in vparts.items():
        out[f"vox_{k}_sha"] = np.array(sha(np.asarray(v).astype(np.int64)))
        out[f"vox_{k}_n"] = np.array(len(v))
    for k, v in mparts.items():
        out[f"mask_{k}"] = np.packbits(v.astype(bool))
    out["mask_keys"] = np.array(list(mparts.keys()))
    vk = ce.extract_top_bottom_voxel_points(vparts)
    ik = ce.extract_top_bottom_image_points(mparts)
    out["vk_keys"], out["vk"] = np.array(list(vk.keys())), np.array([vk[k] for k in vk], dtype=np.float64)
    out["ik_keys"], out["ik"] = np.array(list(ik.keys())), np.array([ik[k] for k in ik], dtype=np.float64)
    vsel, isel = ce.extract_minaret_kps_for_view(grid, front, colours)
    out["sel_keys"] = np.array(list(vsel.keys()))
    out["sel_v"], out["sel_i"] = np.array([vsel[k] for k in vsel], dtype=np.float64), np.array([isel[k] for k in isel], dtype=np.float64)
    for loss in ("L2", "L1"):
        with contextlib.redirect_stdout(buf):
            fit = ce.optimize_camera_with_keypoints(vsel, isel, front, init, loss_type=loss)
        out[f"fit_{loss}"] = np.array([*fit["cam_pos"], *fit["target"], fit["f"], fit["cx"], fit["cy"]], dtype=np.float64)
    rbuf = io.StringIO()                                   # visualize_reprojection's printed table (projection_utils.py:26-67)
    with contextlib.redirect_stdout(rbuf):
        ref.pu.visualize_reprojection(front, vsel, isel, init, title="Front | Initial Reprojection")
    out["reproj_log"] = np.array(rbuf.getvalue())
    np.savez_compressed(os.path.join(HERE, "init_golden.npz"), **out)
    print("init", out["init_row"], out["init_dtypes"], "sel", list(out["sel_keys"]), "fit", out["fit_L2"])


def handoff_golden():
    """voxel_grid_to_points of the live reference (voxel_utils.py:35-51, RGB branch) on the half-resolution Taj grid."""
    grid = np.load(os.path.join(HERE, "aligner_golden.npz"))["grid"]
    out = {}
    for stride in (1, 2, 3):
        pts, cols, shp = ref.vu.voxel_grid_to_points(grid, stride=stride)
        out[f"s{stride}_pts_sha"], out[f"s{stride}_cols_sha"] = np.array(sha(pts)), np.array(sha(cols))
        out[f"s{stride}_n"], out[f"s{stride}_shape"] = np.array(len(pts)), np.array(shp)
        out[f"s{stride}_dtypes"] = np.array([str(pts.dtype), str(cols.dtype)])
    np.savez_compressed(os.path.join(HERE, "handoff_golden.npz"), **out)
    print("handoff", {k: (int(v) if v.ndim == 0 and v.dtype.kind == "i" else None) for k, v in out.items() if k.endswith("_n")})


def build_table_scene(tmp, cams_json):
    """Directory layout notebook 4 expects (nb4 cell 3), filled with the half-resolution Taj scene: the grid of
    aligner_golden.npz, the deformed grid of deform_golden.npz, the original front-mask PNG and the given cameras."""
    ag = np.load(os.path.join(HERE, "aligner_golden.npz"))
    dg = np.load(os.path.join(HERE, "deform_golden.npz"))
    grid = ag["grid"]
    deformed = np.zeros((int(np.prod(grid.shape[:3])), 3), np.uint8)
    deformed[dg["f64_grid_nz"]] = dg["f64_grid_rgb"]
    dirs = {k: os.path.join(tmp, k) for k in ("voxels", "deformed", "cams", "data")}
    for d in dirs.values():
        os.makedirs(d, exist_ok=True)
    os.makedirs(os.path.join(dirs["data"], "Taj", "masks"), exist_ok=True)
    np.savez_compressed(os.path.join(dirs["voxels"], "Taj_voxel_grid.npz"), voxel_grid=grid)
    np.savez_compressed(os.path.join(dirs["deformed"], "Taj_deformed_voxel_grid.npz"), voxel_grid=deformed.reshape(grid.shape))
    shutil.copy(os.path.join(DATA, "Taj", "masks", "Taj_front_mask.png"), os.path.join(dirs["data"], "Taj", "masks"))
    for tag, cam in cams_json.items():
        json.dump(cam, open(os.path.join(dirs["cams"], f"Taj_camera_params_{tag}.json"), "w"))
    return dirs


def tables_golden():
    """The three table drivers of notebook 4 (eval_helpers_intra.py:287-748) run live, visualize=False, on the
    half-resolution Taj scene (the reference's depth buffer is a Python loop over every voxel).  Same two library shims
    as init_golden (ndarray.ptp, skimage label/regionprops)."""
    import tempfile
    import types
    ce = ref.ce

    class _Arr(np.ndarray):
        def ptp(self, *a, **k):
            return np.ptp(np.asarray(self), *a, **k)

    npx = types.ModuleType("numpy_with_ptp")
    npx.__dict__.update(np.__dict__)
    npx.argwhere = lambda m: np.argwhere(m).view(_Arr)
    ce.np = npx

    class _Region:
        def __init__(self, lab, cid):
            yy, xx = np.nonzero(lab == cid)
            self.area, self.centroid, self.label = len(yy), (yy.mean(), xx.mean()), cid

    ce.label2d = lambda mask: scipy.ndimage.label(np.asarray(mask) != 0, structure=np.ones((3, 3), int))[0]
    ce.regionprops = lambda lab: [_Region(lab, c) for c in range(1, int(lab.max()) + 1)]
    import utils.eval_helpers_intra as eh

    cams = {}
    for tag in ("init", "kp", "final"):
        c = json.load(open(os.path.join(ref.root, "results/2.Perspective_Camera_Estimation", f"Taj_camera_params_{tag}.json")))["front"]
        cams[tag] = {"front": {"cam_pos": [v / 2 for v in c["cam_pos"]], "target": [v / 2 for v in c["target"]],
                               "f": c["f"] / 2, "cx": c["cx"] / 2, "cy": c["cy"] / 2}}
    out = {"cams_json": np.array(json.dumps(cams))}
    with tempfile.TemporaryDirectory() as tmp:
        d = build_table_scene(tmp, cams)
        common = dict(monuments=["Taj"], view="front", root_masks=d["data"], cam_dir=d["cams"], part_colors=C.PART_COLORS, visualize=False)
        for name, fn, extra in (("kp", eh.run_minaret_kp_evaluation, {"root_voxels": d["voxels"]}),
                                ("iou", eh.run_minaret_iou_evaluation, {"root_voxels": d["voxels"]}),
                                ("part", eh.run_part_minaret_binary_iou, {"root_voxels": d["voxels"], "deformed_voxels": d["deformed"]})):
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                df = fn(**common, **extra)
            out[f"{name}_df"] = np.array(df.to_json())
            out[f"{name}_log"] = np.array(buf.getvalue())
            print(name, df.to_dict())
    np.savez_compressed(os.path.join(HERE, "tables_golden.npz"), **out)


def depth_golden():
    """compute_global_depth_buffer / project_part_visible of the live utils/eval_helpers_intra.py (:134-190) on the
    half-resolution Taj grid stored in aligner_golden.npz, float32 cameras (as load_camera_json makes them) and float64."""
    import importlib
    eh = importlib.import_module("utils.eval_helpers_intra")
    grid = np.load(os.path.join(HERE, "aligner_golden.npz"))["grid"]
    cams = json.load(open(os.path.join(ref.root, "results/2.Perspective_Camera_Estimation/Taj_camera_params_final.json")))
    H, W = 139, 256
    out = {}
    for view in ("front", "drone"):
        c = cams[view]
        for dt in (np.float32, np.float64):
            cam = {"cam_pos": (np.array(c["cam_pos"]) / 2).astype(dt), "target": (np.array(c["target"]) / 2).astype(dt),
                   "f": c["f"] / 2, "cx": c["cx"] / 2, "cy": c["cy"] / 2}
            key = f"{view}_{np.dtype(dt).name}"
            z = eh.compute_global_depth_buffer(grid, cam, H, W)
            out[key + "_cam"] = np.array([*cam["cam_pos"], *cam["target"], cam["f"], cam["cx"], cam["cy"]], dtype=np.float64)
            out[key + "_zbuf"] = z
            for tag, parts in (("min", ["front_minarets", "back_minarets"]), ("dome", ["dome"])):
                pts, _ = ref.vu.get_voxel_points_by_parts(grid, C.PART_COLORS, parts)
                out[f"{key}_{tag}_visible"] = np.packbits(eh.project_part_visible(pts, cam, z, H, W))
                out[f"{key}_{tag}_loose"] = np.packbits(eh.project_part_visible(pts, cam, z, H, W, eps=0.75))
            print(key, "finite", int(np.isfinite(z).sum()))
    np.savez_compressed(os.path.join(HERE, "depth_golden.npz"), **out)


def real5_golden():
    """The two monuments carve_golden.npz does not hold (SURVEY 4 asks for all five): Charminar -- a PORTRAIT mask (width
    177 at max_dim 256, 88 at 128: no multiple of 32, the branch of mask_utils.py:65-71) -- and Itimad, through the live
    global_carve + partwise_carve with notebook 1's jobs.  Masks, hashes, occupied counts and the printed component log;
    full arrays only at max_dim 128."""
    out = {}
    cases = [("Charminar", 128), ("Charminar", 256), ("Itimad", 256)]
    for name, md in cases:
        sem, ext, binm = ref.mu.load_and_prepare_masks(os.path.join(ref.root, "data"), name, "front", md,
                                                       C.PART_COLORS_NP, C.INTERIOR_PARTS)
        g = ref.vc.global_carve(binm, ext, 90)
        pc = ref.vc.part_carve(g, ext, GROUP_JOBS)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            p = ref.vc.partwise_carve(g, ext, sem, C.PART_COLORS_NP, GROUP_JOBS, PART_SYMMETRY, EXTRUSION_DEPTHS)
        key = f"real_{name}_{md}"
        out[key + "_sem"], out[key + "_ext"], out[key + "_bin"] = sem, ext, binm
        out[key + "_global_sha"], out[key + "_partcarve_sha"], out[key + "_partwise_sha"] = np.array(sha(g)), np.array(sha(pc)), np.array(sha(p))
        out[key + "_global_occ"] = np.array(np.count_nonzero(g.any(-1)))
        out[key + "_partwise_occ"] = np.array(np.count_nonzero(p.any(-1)))
        out[key + "_log"] = np.array(buf.getvalue())
        if md <= 128:
            out[key + "_global"], out[key + "_partwise"] = g, p
        print(key, sem.shape, g.shape, out[key + "_partwise_sha"], out[key + "_partwise_occ"])
    out["real_cases"] = np.array([f"{n}_{m}" for n, m in cases])
    np.savez_compressed(os.path.join(HERE, "real5_golden.npz"), **out)


def topk_golden():
    """extract_top_k_components of the live reference (voxel_utils.py:22-31: 26-connectivity, np.ptp along axis 1, stable
    sort) on small random grids, incl. corner-touching blobs and height ties."""
    import utils.voxel_utils as rvu
    rng = np.random.default_rng(SEED + 31)
    out = {}
    colour = np.array(C.PART_COLORS["front_minarets"])
    cases = []
    for i, (shape, p) in enumerate([((14, 20, 12), 0.2), ((9, 30, 9), 0.3), ((12, 12, 12), 0.07), ((6, 25, 6), 0.45)]):
        grid = np.zeros(shape + (3,), np.uint8)
        grid[rng.random(shape) < p] = colour
        grid[rng.random(shape) < 0.1] = C.PART_COLORS["dome"]
        out[f"g{i}"] = grid
        for k in (1, 2, 4):
            out[f"g{i}_k{k}"] = rvu.extract_top_k_components(grid, colour, k)
        cases.append(i)
    out["n"] = np.array(len(cases))
    out["colour"] = colour
    np.savez_compressed(os.path.join(HERE, "topk_golden.npz"), **out)


def partcarve_asym_golden():
    """part_carve (voxel_carving_utils.py:139-160) of grids that are NOT 4-way symmetric: the inputs on which the rotated
    source occupancy decides (global_carve's output never is).  Random sparse grids coloured column-wise from a blocky
    semantic image, every notebook group at 90 degrees, through the live reference."""
    out = {}
    rng = np.random.default_rng(SEED + 17)
    cases = [("w32h8", 32, 8, 0.55), ("w64h7", 64, 7, 0.6), ("w32h32", 32, 32, 0.5), ("w96h6", 96, 6, 0.65)]
    for tag, W, H, p in cases:
        sem = blocky_semantic(rng, H, W, False)
        ext = sem.copy()
        for q in C.INTERIOR_PARTS:
            ext[np.all(sem == C.PART_COLORS_NP[q], axis=-1)] = C.PART_COLORS_NP["full_building"]
        occ = rng.random((W, H, W)) < p
        grid = np.zeros((W, H, W, 3), np.uint8)
        grid[occ] = ext.transpose(1, 0, 2)[:, :, None, :].repeat(W, axis=2)[occ]
        bg = np.all(grid == C.PART_COLORS_NP["background"], axis=-1)
        grid[bg] = 0                                                   # background columns carry no voxels
        pc = ref.vc.part_carve(grid, ext, GROUP_JOBS)
        key = "asym_" + tag
        out[key + "_ext"], out[key + "_grid"], out[key + "_partcarve"] = ext, grid, pc
        n_in, n_out = np.count_nonzero(grid.any(-1)), np.count_nonzero(pc.any(-1))
        assert 0 < n_out < n_in
        print(key, grid.shape, n_in, n_out)
    out["asym_cases"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(HERE, "partcarve_asym_golden.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["assets", "camera", "carve", "aligner", "depth", "deform", "init", "handoff", "tables", "partcarve_asym"]
    if "assets" in which:
        copy_assets()
    if "camera" in which:
        camera_golden()
    if "carve" in which:
        carve_golden()
    if "aligner" in which:
        aligner_golden()
    if "depth" in which:
        depth_golden()
    if "deform" in which:
        deform_golden()
    if "init" in which:
        init_golden()
    if "handoff" in which:
        handoff_golden()
    if "tables" in which:
        tables_golden()
    if "partcarve_asym" in which:
        partcarve_asym_golden()
    if "real5" in which:
        real5_golden()
    if "topk" in which:
        topk_golden()
    for f in ("camera_golden.npz", "carve_golden.npz", "aligner_golden.npz", "depth_golden.npz", "deform_golden.npz", "init_golden.npz", "handoff_golden.npz", "tables_golden.npz", "partcarve_asym_golden.npz", "real5_golden.npz", "topk_golden.npz"):
        if os.path.exists(os.path.join(HERE, f)):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
