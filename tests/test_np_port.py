"""The NumPy port used as the timed CPU baseline (oracle/np_port.py) against the C oracle, so that the
baseline bench.py reports is known to compute the same thing."""
import numpy as np

from conftest import pkg


def test_np_port_matches_oracle(oracle):
    from oracle import np_port
    syn = pkg("synthetic")
    N, H, W = 64, 192, 160
    rgb = syn.label_lut()[syn.monument_labels(N).numpy()]
    parts = syn.PART_NAMES
    pts, cols = oracle.get_voxel_points_by_parts(rgb, oracle.PART_COLORS, parts)
    base = syn.base_camera(N, H, W)
    gt = oracle.project_colored_voxels(pts, cols, base[0:3] + 2.0, base[3:6], base[6], base[7], base[8], H, W)
    seg = oracle.mask_parts_from_image(gt, oracle.PART_COLORS, parts)
    sel = {p: oracle.PART_COLORS[p] for p in parts}
    cand = syn.candidates(base, 6)
    for row in cand:
        img = np_port.render(pts, cols, row[0:3], row[3:6], row[6], row[7], row[8], H, W)
        assert np.array_equal(img, oracle.project_colored_voxels(pts, cols, row[0:3], row[3:6], row[6], row[7], row[8], H, W))
        counts, mean = np_port.partwise_iou(img, seg, sel)
        inter, uni = oracle.partwise_counts(img, seg, sel)
        assert counts == list(zip(inter.tolist(), uni.tolist()))
        assert mean == oracle.compute_partwise_iou(img, seg, sel)[1]
    dt1, s1 = np_port.timed_sweep(pts, cols, seg, sel, cand, H, W, processes=1)
    dt2, s2 = np_port.timed_sweep(pts, cols, seg, sel, cand, H, W, processes=2)
    assert s1 == s2 and dt1 > 0 and dt2 > 0
