"""GPU parity tests of the camera-scoring path: CUDA kernels (through the C ABI and the
reference-signature Python layer) against the oracle and the golden vectors."""
import warnings

import numpy as np
import pytest
import torch

from conftest import pkg
from helpers import row_to_args, sha

pytestmark = pytest.mark.gpu

MINARETS = ["front_minarets", "back_minarets"]


@pytest.fixture(scope="module")
def mods():
    class M:
        pass
    m = M()
    m.nv = pkg("utils._native")
    m.eng = pkg("utils._engine")
    m.cg = pkg("utils.camera_geometry")
    m.pu = pkg("utils.projection_utils")
    m.vu = pkg("utils.voxel_utils")
    m.mu = pkg("utils.mask_utils")
    m.ce = pkg("utils.camera_estimation")
    m.cfg = pkg("utils.config")
    m.syn = pkg("synthetic")
    assert torch.cuda.is_available()
    return m


def all_parts(cfg):
    return [p for p in cfg.PART_COLORS if p != "background"]


def test_look_at_batch_bit_exact(mods, camera_golden):
    g = camera_golden
    K = g["lookat_eye"].shape[0]
    cand = np.zeros((K, 9))
    cand[:, 0:3], cand[:, 3:6] = g["lookat_eye"], g["lookat_target"]
    cams = mods.eng.setup_cameras(torch.from_numpy(cand).cuda()).cpu().numpy()
    assert np.array_equal(cams[:, 3:12].reshape(K, 3, 3), g["lookat_R64"])
    cams32 = mods.eng.setup_cameras(torch.from_numpy(cand.astype(np.float32)).cuda()).cpu().numpy()
    assert np.array_equal(cams32[:, 3:12].reshape(K, 3, 3), g["lookat_R32"])
    R = mods.cg.look_at_rotation(g["lookat_eye"][3], g["lookat_target"][3])
    assert R.dtype == np.float64 and np.array_equal(R, g["lookat_R64"][3])


def test_points_by_parts_matches_oracle(mods, oracle, taj, camera_golden):
    for parts in (MINARETS, all_parts(mods.cfg), ["windows"], []):
        pts, cols = mods.vu.get_voxel_points_by_parts(taj["grid"], mods.cfg.PART_COLORS, parts)
        opts, ocols = oracle.get_voxel_points_by_parts(taj["grid"], oracle.PART_COLORS, parts)
        assert pts.dtype == np.float32 and cols.dtype == np.uint8
        assert np.array_equal(pts, opts) and np.array_equal(cols, ocols)
    g = camera_golden["adv_grid"]          # ragged tiny grid (not a multiple of the tile size)
    pts, cols = mods.vu.get_voxel_points_by_parts(g, mods.cfg.PART_COLORS, ["dome", "plinth", "windows"])
    opts, ocols = oracle.get_voxel_points_by_parts(g, oracle.PART_COLORS, ["dome", "plinth", "windows"])
    assert np.array_equal(pts, opts) and np.array_equal(cols, ocols)


def test_mask_parts_from_image(mods, oracle, taj):
    for view in ("front", "drone"):
        got = mods.mu.mask_parts_from_image(taj[view], mods.cfg.PART_COLORS, MINARETS)
        assert np.array_equal(got, oracle.mask_parts_from_image(taj[view], oracle.PART_COLORS, MINARETS))


@pytest.mark.parametrize("view", ["front", "drone"])
@pytest.mark.parametrize("tag", ["min", "all"])
def test_project_colored_voxels_images(mods, oracle, camera_golden, taj, view, tag):
    g = camera_golden
    parts = MINARETS if tag == "min" else all_parts(mods.cfg)
    H, W = taj[view].shape[:2]
    pts, cols = oracle.get_voxel_points_by_parts(taj["grid"], oracle.PART_COLORS, parts)
    seg = oracle.mask_parts_from_image(taj[view], oracle.PART_COLORS, parts)
    sel = {p: mods.cfg.PART_COLORS[p] for p in parts}
    for k, row in enumerate(g[f"taj_{view}_{tag}_cand"][:6]):
        img = mods.pu.project_colored_voxels(pts, cols, *row_to_args(row), H, W)
        assert img.shape == (H, W, 3) and img.dtype == np.uint8
        assert sha(img) == str(g[f"taj_{view}_{tag}_sha"][k])
        per, mean = mods.ce.compute_partwise_iou(img, seg, sel)
        assert mean == g[f"taj_{view}_{tag}_scores"][k]
        operp, omean = oracle.compute_partwise_iou(img, seg, sel)
        assert per == operp and mean == omean


@pytest.mark.parametrize("view", ["front", "drone"])
@pytest.mark.parametrize("tag", ["min", "all"])
def test_sweep_counts_and_scores(mods, camera_golden, taj, view, tag):
    g = camera_golden
    parts = MINARETS if tag == "min" else all_parts(mods.cfg)
    cand = g[f"taj_{view}_{tag}_cand"]
    scores, counts, best = mods.ce.score_camera_candidates(taj["grid"], taj[view], mods.cfg.PART_COLORS, parts, cand)
    assert counts.dtype == np.int64 and np.array_equal(counts, g[f"taj_{view}_{tag}_counts"])
    assert np.array_equal(scores, g[f"taj_{view}_{tag}_scores"])          # identical, not just 1e-5
    assert best == int(np.argmax(g[f"taj_{view}_{tag}_scores"]))


@pytest.mark.parametrize("view", ["front", "drone"])
def test_sweep_float32(mods, camera_golden, taj, view):
    g = camera_golden
    scores, counts, _ = mods.ce.score_camera_candidates(taj["grid"], taj[view], mods.cfg.PART_COLORS, MINARETS,
                                                        g[f"taj_{view}_f32_cand"], dtype=np.float32)
    assert np.array_equal(counts, g[f"taj_{view}_f32_counts"])


def test_projection_float32_images(mods, oracle, camera_golden, taj):
    g = camera_golden
    H, W = taj["front"].shape[:2]
    pts, cols = oracle.get_voxel_points_by_parts(taj["grid"], oracle.PART_COLORS, MINARETS)
    for k, row in enumerate(g["taj_front_f32_cand"]):
        img = mods.pu.project_colored_voxels(pts, cols, *row_to_args(row, np.float32), H, W)
        assert sha(img) == str(g["taj_front_f32_sha"][k])


def test_adversarial_cameras(mods, oracle, camera_golden):
    """Points behind the camera (clamped, not culled), straight-down view, eye == target (NaN), exact .5 ties,
    a part absent from both grid and image."""
    g = camera_golden
    parts = ["dome", "plinth", "windows"]
    scores, counts, best = mods.ce.score_camera_candidates(g["adv_grid"], g["adv_image"], mods.cfg.PART_COLORS, parts,
                                                           g["adv_cand"])
    assert np.array_equal(counts, g["adv_counts"])
    assert np.array_equal(scores, g["adv_scores"])
    pts, cols = oracle.get_voxel_points_by_parts(g["adv_grid"], oracle.PART_COLORS, parts)
    for k, row in enumerate(g["adv_cand"]):
        img = mods.pu.project_colored_voxels(pts, cols, *row_to_args(row), 20, 24)
        assert sha(img) == str(g["adv_sha"][k]), k


def test_per_part_mode(mods, camera_golden, taj):
    g = camera_golden
    parts = [str(p) for p in g["taj_front_perpart_parts"]]
    p = taj["cams"]["front"]
    row = np.array([*p["cam_pos"], *p["target"], p["f"], p["cx"], p["cy"]])
    scorer = mods.ce.CandidateScorer(taj["grid"], taj["front"], mods.cfg.PART_COLORS, parts, mode="per_part")
    scores, counts, _ = scorer.score(np.stack([row, row]))
    assert counts.shape == (2, len(parts) + 1, 2)
    assert np.array_equal(counts[0], g["taj_front_perpart_counts"])
    assert np.array_equal(counts[1], counts[0])


def test_set_image_switches_the_view_without_a_rebuild(mods, taj):
    """CandidateScorer.set_image: the drone mask scored by a scorer built on the front mask equals a scorer built on the
    drone mask (joint and per-part mode, images of different sizes), and switching back restores the front scores."""
    cams = taj["cams"]
    rows = {v: np.array([*cams[v]["cam_pos"], *cams[v]["target"], cams[v]["f"], cams[v]["cx"], cams[v]["cy"]]) for v in ("front", "drone")}
    cand = {v: mods.ce.random_candidates(rows[v], 9, np.random.default_rng(2)) for v in rows}
    for mode in ("joint", "per_part"):
        a = mods.ce.CandidateScorer(taj["grid"], taj["front"], mods.cfg.PART_COLORS, MINARETS, mode=mode)
        f0 = a.score(cand["front"])
        a.set_image(taj["drone"])
        assert (a.H, a.W) == taj["drone"].shape[:2]
        d0 = a.score(cand["drone"])
        b = mods.ce.CandidateScorer(taj["grid"], taj["drone"], mods.cfg.PART_COLORS, MINARETS, mode=mode)
        d1 = b.score(cand["drone"])
        assert np.array_equal(d0[0], d1[0]) and np.array_equal(d0[1], d1[1]) and d0[2] == d1[2]
        f1 = a.set_image(taj["front"]).score(cand["front"])
        assert np.array_equal(f0[0], f1[0]) and np.array_equal(f0[1], f1[1])


def test_batching_determinism_and_ties(mods, taj):
    """K larger than one z-buffer batch; repeated runs byte-identical; duplicate best -> first index."""
    p = taj["cams"]["front"]
    base = np.array([*p["cam_pos"], *p["target"], p["f"], p["cx"], p["cy"]])
    cand = mods.ce.random_candidates(base, 150, np.random.default_rng(5))
    cand[70] = cand[0]
    cand[149] = cand[0]
    scorer = mods.ce.CandidateScorer(taj["grid"], taj["front"], mods.cfg.PART_COLORS, MINARETS)
    s1, c1, b1 = scorer.score(cand)
    s2, c2, b2 = scorer.score(cand)
    assert np.array_equal(s1, s2) and np.array_equal(c1, c2) and b1 == b2
    assert np.array_equal(c1[0], c1[70]) and np.array_equal(c1[0], c1[149])
    assert b1 == int(np.argmax(s1))
    one, cone, _ = scorer.score(cand[37:38])
    assert one[0] == s1[37] and np.array_equal(cone[0], c1[37])


def test_synthetic_monument_sweep_vs_oracle(mods, oracle):
    """Config-3-shaped case at a size the oracle finishes in seconds: 128^3 grid, 1024^2 mask, all parts."""
    N, H, W, K = 128, 1024, 1024, 12
    syn = mods.syn
    labels = syn.monument_labels(N)
    rgb = syn.label_lut()[labels.numpy()]
    parts = syn.PART_NAMES
    base = syn.base_camera(N, H, W, "front")
    pts, cols = oracle.get_voxel_points_by_parts(rgb, oracle.PART_COLORS, parts)
    hidden = base + np.array([3.0, -2.0, 5.0, 1.0, 2.0, -3.0, 4.0, 1.5, -2.5])
    gt = oracle.project_colored_voxels(pts, cols, hidden[0:3], hidden[3:6], hidden[6], hidden[7], hidden[8], H, W)
    cand = syn.candidates(base, K)
    scores, counts, best = mods.ce.score_camera_candidates(rgb, gt, mods.cfg.PART_COLORS, parts, cand)
    sel = {p: oracle.PART_COLORS[p] for p in parts}
    seg = oracle.mask_parts_from_image(gt, oracle.PART_COLORS, parts)
    for k in range(K):
        s, inter, uni = oracle.score_candidate(pts, cols, seg, sel,
                                               {"cam_pos": cand[k, 0:3], "target": cand[k, 3:6], "f": cand[k, 6],
                                                "cx": cand[k, 7], "cy": cand[k, 8]}, H, W)
        assert np.array_equal(counts[k, :, 0], inter) and np.array_equal(counts[k, :, 1], uni), k
        assert scores[k] == s
    # device-resident label grid path gives the same answer
    scorer = mods.ce.CandidateScorer(torch.from_numpy(rgb).cuda(), torch.from_numpy(gt).cuda(), mods.cfg.PART_COLORS, parts)
    s2, c2, b2 = scorer.score(cand)
    assert np.array_equal(s2, scores) and np.array_equal(c2, counts) and b2 == best


def test_numpy_mean_order_on_device(mods):
    """scores use NumPy's pairwise summation order for every part count 1..32."""
    rng = np.random.default_rng(11)
    for P in list(range(1, 13)) + [16, 17, 25, 32]:
        # P parts, each a 1-pixel-wide stripe with random intersection/union sizes
        H, W = 64, 2 * P
        names = [f"p{i}" for i in range(P)]
        colors = {n: (10 + i, 200 - i, 3 * i + 1) for i, n in enumerate(names)}
        img = np.zeros((H, W, 3), np.uint8)
        grid = np.zeros((1, H, W, 3), np.uint8)
        for i, n in enumerate(names):
            a, b = rng.integers(1, H, 2)
            img[:a, 2 * i] = colors[n]
            grid[0, :b, 2 * i] = colors[n]
        # orthographic-like camera far away looking down +z so that voxel (x, y) lands on pixel (x, H-1-y)
        cand = np.array([[W / 2 - 0.5, H / 2 - 0.5, -1e6, W / 2 - 0.5, H / 2 - 0.5, 0.0, 1e6, W / 2 - 0.5, H / 2 - 0.5]])
        scores, counts, _ = mods.ce.score_camera_candidates(grid, img, colors, names, cand)
        ious = [c[0] / c[1] if c[1] > 0 else 0.0 for c in counts[0]]
        assert scores[0] == np.mean(ious), P


def test_filtered_splat_is_bit_identical_to_exact_fp64(mods, taj, monkeypatch):
    """The FP32-filter + FP64-queue kernel against the plain FP64 kernel: identical z-buffers, including
    cameras inside the cloud, behind it, degenerate (eye == target), tiny / huge focal lengths and far offsets."""
    rng = np.random.default_rng(77)
    dev = torch.device("cuda")
    p = taj["cams"]["front"]
    base = np.array([*p["cam_pos"], *p["target"], p["f"], p["cx"], p["cy"]])
    cand = mods.ce.random_candidates(base, 24, rng)
    wild = np.tile(base, (12, 1))
    wild[0, 0:3] = [250.0, 100.0, 250.0]                       # inside the model
    wild[1, 0:3] = wild[1, 3:6]                                # eye == target
    wild[2, 6] = 1e-2                                          # tiny f
    wild[3, 6] = 1e5                                           # huge f
    wild[4, 2] = 2000.0                                        # behind, looking back
    wild[5, 0:3] = [1e6, -3e5, -2e6]                           # very far
    wild[6, 6] = -500.0                                        # negative f
    wild[7, 7:9] = [-5000.0, 9000.0]                           # principal point far off
    wild[8, 0:3] = [256.3, 139.2, 255.9]                       # inside, off-grid
    wild[9, 0:3] = wild[9, 3:6] + [0.0, 300.0, 0.0]            # straight down
    wild[10, 6] = float("nan")
    wild[11, 0] = float("inf")
    cand = np.concatenate([cand, wild])
    parts = all_parts(mods.cfg)
    pts, pt_label, _, _ = mods.vu.device_points_by_parts(taj["grid"], mods.cfg.PART_COLORS, parts)
    for (H, W) in ((278, 512), (64, 48)):
        cams = mods.eng.setup_cameras(torch.from_numpy(cand).to(dev))
        for mode in (mods.nv.MODE_JOINT, mods.nv.MODE_PER_PART):
            monkeypatch.setenv("P3D_SPLAT_EXACT", "1")
            ref = mods.eng.splat(pts, pt_label, cams, H, W, mode).cpu()
            monkeypatch.setenv("P3D_SPLAT_EXACT", "0")
            got = mods.eng.splat(pts, pt_label, cams, H, W, mode).cpu()
            assert torch.equal(ref, got), (H, W, mode, int((ref != got).sum()))
    # synthetic 256^3, 1024^2 image: the benchmark geometry
    syn = mods.syn
    rgb = torch.from_numpy(syn.label_lut()).to(dev)[syn.monument_labels(256, dev).long()]
    pts, pt_label, _, _ = mods.vu.device_points_by_parts(rgb, mods.cfg.PART_COLORS, syn.PART_NAMES)
    cand = syn.candidates(syn.base_camera(256, 1024, 1024), 48)
    cams = mods.eng.setup_cameras(torch.from_numpy(cand).to(dev))
    monkeypatch.setenv("P3D_SPLAT_EXACT", "1")
    ref = mods.eng.splat(pts, pt_label, cams, 1024, 1024, mods.nv.MODE_JOINT).cpu()
    monkeypatch.setenv("P3D_SPLAT_EXACT", "0")
    got = mods.eng.splat(pts, pt_label, cams, 1024, 1024, mods.nv.MODE_JOINT).cpu()
    assert torch.equal(ref, got)


def test_randomised_sweeps_vs_oracle(mods, oracle):
    """Differential test over the rarely exercised corners of the sweep: odd image sizes (scalar score path), one to
    nine parts, parts without voxels or without pixels, tiny point lists, K from 1 to 300 (several z-buffer batches and
    camera groups), float64 and float32 candidates -- counts and scores against the oracle."""
    ce = pkg("utils.camera_estimation")
    rng = np.random.default_rng(2024)
    names = [k for k in oracle.PART_COLORS if k != "background"]
    for trial in range(14):
        A0, A1, A2 = (int(v) for v in rng.integers(3, 22, 3))
        H, W = (int(v) for v in rng.integers(5, 70, 2))
        parts = list(rng.choice(names, size=int(rng.integers(1, 10)), replace=False))
        grid = np.zeros((A0, A1, A2, 3), np.uint8)
        lab = rng.integers(0, len(parts) + 2, (A0, A1, A2))
        for k, n in enumerate(parts[:-1] if len(parts) > 2 else parts):        # sometimes a selected part has no voxels
            grid[lab == k + 1] = oracle.PART_COLORS[n]
        image = np.zeros((H, W, 3), np.uint8)
        ilab = rng.integers(0, len(parts) + 1, (H, W))
        for k, n in enumerate(parts[1:] if len(parts) > 2 else parts):        # ... or no pixels in the image
            image[ilab == k + 1] = oracle.PART_COLORS[n]
        K = int(rng.choice([1, 2, 7, 33, 130, 300]))
        dt = np.float32 if trial % 3 == 2 else np.float64
        ctr = np.array([A2, A1, A0]) / 2
        cand = np.empty((K, 9))
        cand[:, 0:3] = ctr + rng.normal(0, 1, (K, 3)) * 3 + np.array([0, 0, -3.0 * max(A0, A1, A2)])
        cand[:, 3:6] = ctr + rng.normal(0, 1, (K, 3))
        cand[:, 6] = rng.uniform(0.5, 3.0, K) * max(H, W)
        cand[:, 7] = W / 2 + rng.normal(0, 3, K)
        cand[:, 8] = H / 2 + rng.normal(0, 3, K)
        cand = cand.astype(dt)
        scorer = ce.CandidateScorer(grid, image, oracle.PART_COLORS, parts, dtype=dt)
        scores, counts, best = scorer.score(cand)
        pts, cols = oracle.get_voxel_points_by_parts(grid, oracle.PART_COLORS, parts)
        seg = oracle.mask_parts_from_image(image, oracle.PART_COLORS, parts)
        sel = {p: oracle.PART_COLORS[p] for p in parts}
        ref_scores = []
        for k in range(K):
            cp, tg, f, cx, cy = row_to_args(cand[k], dt)
            s, inter, uni = oracle.score_candidate(pts, cols, seg, sel, {"cam_pos": cp, "target": tg, "f": f, "cx": cx, "cy": cy}, H, W)
            ref_scores.append(s)
            assert np.array_equal(counts[k, :, 0], inter) and np.array_equal(counts[k, :, 1], uni), (trial, k)
            assert scores[k] == s, (trial, k)
        assert best == int(np.argmax(ref_scores))
