"""The segment splat (one thread per x-run of the point list, csrc/p3d_camera.cu) against the per-point splat and the
oracle: the segment records themselves, then bit-identical counts/scores in every mode and dtype for cameras that see
the whole object (no bounds test in the kernel), part of it, none of it, and cameras inside the cloud."""
import numpy as np
import pytest
import torch

from conftest import pkg
from helpers import row_to_args

pytestmark = pytest.mark.gpu


def blocky_grid(rng, shape, parts, colors, n_boxes=14):
    """Solid boxes of random parts: long x-runs, label changes inside rows, rows that start at odd x."""
    A0, A1, A2 = shape
    grid = np.zeros((A0, A1, A2, 3), np.uint8)
    for _ in range(n_boxes):
        lo = [int(rng.integers(0, s - 2)) for s in shape]
        hi = [int(rng.integers(l + 1, s + 1)) for l, s in zip(lo, shape)]
        grid[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = colors[parts[int(rng.integers(0, len(parts)))]]
    return grid


def cameras(rng, shape, H, W, K, dt):
    A0, A1, A2 = shape
    ctr = np.array([A2, A1, A0]) / 2
    size = float(max(shape))
    cand = np.empty((K, 9))
    dist = rng.choice([0.3, 2.5, 2.5, 4.0, 12.0], K) * size                 # 0.3: inside / next to the cloud
    dirs = rng.normal(0, 1, (K, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    cand[:, 0:3] = ctr + dirs * dist[:, None]
    cand[:, 3:6] = ctr + rng.normal(0, 1, (K, 3)) * size * 0.05
    cand[:, 6] = rng.uniform(0.6, 1.5, K) * max(H, W) * (dist / size) / 2.0   # object fills roughly 1/3 .. 3/4 of the image
    off = rng.choice([0.0, 0.0, 0.0, 0.45, -0.45, 1.5], K)
    cand[:, 7] = W / 2 + off * W + rng.normal(0, 2, K)
    cand[:, 8] = H / 2 + rng.choice([0.0, 0.0, 0.4, -1.5], K) * H + rng.normal(0, 2, K)
    return cand.astype(dt)


def test_segment_records_match_the_numpy_restatement(oracle):
    eng, nv = pkg("utils._engine"), pkg("utils._native")
    rng = np.random.default_rng(5)
    L = int(nv.lib.p3d_segment_length())
    names = [k for k in oracle.PART_COLORS if k != "background"]
    for shape in ((9, 7, 40), (5, 6, 129), (3, 4, 600), (3, 3, 8), (2, 2, 1)):
        parts = names[:5]
        for grid in (blocky_grid(rng, shape, parts, oracle.PART_COLORS) if min(shape) > 2 else np.zeros(shape + (3,), np.uint8),
                     np.where(rng.random(shape + (1,)) < 0.6, np.array(oracle.PART_COLORS[parts[0]], np.uint8), 0).astype(np.uint8)):
            pts, cols = oracle.get_voxel_points_by_parts(grid, oracle.PART_COLORS, parts)
            lut = {tuple(oracle.PART_COLORS[p]): i + 1 for i, p in enumerate(parts)}
            lab = np.array([lut[tuple(c)] for c in cols], np.uint8)
            want, bad = oracle.point_segments(pts, lab, L)
            assert bad == 0
            got = eng.build_segments(torch.from_numpy(pts).cuda().contiguous(), torch.from_numpy(lab).cuda())
            if len(pts) == 0:
                assert got is None
                continue
            assert np.array_equal(got.cpu().numpy().view(np.uint32), want)
            lens = ((want[:, 1] >> 16) & 0xf) + 1
            assert lens.sum() == len(pts) and lens.max() <= L
            # every point is owned by exactly one (segment, j)
            Ts = ((want[:, 1] >> 20) & 0x3f) + 1
            owned = np.concatenate([f + t * np.arange(c) for f, t, c in zip(want[:, 2].astype(np.int64), Ts, lens)])
            assert np.array_equal(np.sort(owned), np.arange(len(pts)))
    # lists the segment form cannot represent are refused (the sweep then runs the per-point splat)
    pts = torch.tensor([[0.5, 1, 1], [1.5, 1, 1]], dtype=torch.float32).cuda()
    assert eng.build_segments(pts, torch.ones(2, dtype=torch.uint8).cuda()) is None
    pts = torch.tensor([[70000.0, 1, 1]], dtype=torch.float32).cuda()
    assert eng.build_segments(pts, torch.ones(1, dtype=torch.uint8).cuda()) is None


@pytest.mark.parametrize("mode", ["joint", "per_part"])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_segment_splat_equals_point_splat_and_oracle(oracle, mode, dt):
    ce = pkg("utils.camera_estimation")
    rng = np.random.default_rng(21 + (dt == np.float32) + 2 * (mode == "joint"))
    names = [k for k in oracle.PART_COLORS if k != "background"]
    inview_seen = 0
    for trial, (shape, (H, W)) in enumerate((((24, 18, 37), (96, 128)), ((16, 30, 21), (75, 90)), ((31, 9, 64), (128, 100)))):
        parts = list(rng.choice(names, size=5, replace=False))
        grid = blocky_grid(rng, shape, parts, oracle.PART_COLORS)
        image = np.zeros((H, W, 3), np.uint8)
        ilab = rng.integers(0, 7, (H // 4 + 1, W // 4 + 1)).repeat(4, 0).repeat(4, 1)[:H, :W]
        for k, n in enumerate(parts):
            image[ilab == k + 1] = oracle.PART_COLORS[n]
        K = 48
        cand = cameras(rng, shape, H, W, K, dt)
        seg_scorer = ce.CandidateScorer(grid, image, oracle.PART_COLORS, parts, dtype=dt, mode=mode, use_segments=True)
        pt_scorer = ce.CandidateScorer(grid, image, oracle.PART_COLORS, parts, dtype=dt, mode=mode, use_segments=False)
        assert seg_scorer.segs is not None and pt_scorer.segs is None
        s1, c1, b1 = seg_scorer.score(cand)
        s2, c2, b2 = pt_scorer.score(cand)
        assert np.array_equal(c1, c2) and np.array_equal(s1, s2) and b1 == b2
        if mode == "joint":
            pts, cols = oracle.get_voxel_points_by_parts(grid, oracle.PART_COLORS, parts)
            seg = oracle.mask_parts_from_image(image, oracle.PART_COLORS, parts)
            sel = {p: oracle.PART_COLORS[p] for p in parts}
            for k in range(K):
                cp, tg, f, cx, cy = row_to_args(cand[k], dt)
                s, inter, uni = oracle.score_candidate(pts, cols, seg, sel, {"cam_pos": cp, "target": tg, "f": f, "cx": cx, "cy": cy}, H, W)
                assert np.array_equal(c1[k, :, 0], inter) and np.array_equal(c1[k, :, 1], uni) and s1[k] == s, (trial, k)
                inview_seen += int(inter.sum() > 0)
    if mode == "joint":
        assert inview_seen > 20                              # the camera mix does put the object into the image
