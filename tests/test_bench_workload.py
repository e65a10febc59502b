"""bench_workload.py (the NumPy-only scene definition the reference arm of bench.py uses) against the package's own
synthetic.py (what the GPU arm builds on the device): identical grids, cameras, candidates and palette; and the CPU arm
loads the reference functions from baseline/_ref when they are installed."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, pkg

sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))


def test_numpy_scene_equals_package_scene():
    import bench_workload as bw
    syn, cfg = pkg("synthetic"), pkg("utils.config")
    assert bw.PART_COLORS == {k: tuple(v) for k, v in cfg.PART_COLORS.items()} and list(bw.PART_COLORS) == list(cfg.PART_COLORS)
    assert bw.PART_NAMES == syn.PART_NAMES and bw.LABEL == syn.LABEL
    assert np.array_equal(bw.label_lut(), syn.label_lut())
    for N in (32, 64, 128):
        assert np.array_equal(bw.monument_labels(N), syn.monument_labels(N).numpy())
    for view in ("front", "aerial"):
        assert np.array_equal(bw.base_camera(512, 1024, 1024, view), syn.base_camera(512, 1024, 1024, view))
    base = bw.base_camera(512, 1024, 1024)
    assert np.array_equal(bw.candidates(base, 257), syn.candidates(base, 257))
    assert bw.CANDIDATE_SEED == syn.CANDIDATE_SEED


def test_points_of_matches_oracle(oracle):
    import bench_workload as bw
    lab = bw.monument_labels(32)
    rgb = bw.label_lut()[lab]
    for names in (bw.PART_NAMES, ["front_minarets", "back_minarets"]):
        pts, cols = bw.points_of(lab, names)
        opts, ocols = oracle.get_voxel_points_by_parts(rgb, oracle.PART_COLORS, names)
        assert np.array_equal(pts, opts) and np.array_equal(cols, ocols)


def test_cpu_arm_matches_oracle(oracle):
    """The CPU arm (reference functions from baseline/_ref when installed, else the NumPy port) gives the oracle's counts."""
    import bench_workload as bw
    import cpu_arm
    N, H, W = 32, 96, 80
    lab = bw.monument_labels(N)
    pts, cols = bw.points_of(lab, bw.PART_NAMES)
    base = bw.base_camera(N, H, W)
    cpu = cpu_arm.CpuScorer(pts, cols, np.zeros((H, W, 3), np.uint8), bw.PART_NAMES, part_colors=bw.PART_COLORS)
    gt = cpu.render(pts, cols, base + bw.HIDDEN_DELTA * 0.1)
    assert np.array_equal(gt, oracle.project_colored_voxels(pts, cols, *(lambda r: (r[0:3], r[3:6], r[6], r[7], r[8]))(base + bw.HIDDEN_DELTA * 0.1), H, W))
    cpu = cpu_arm.CpuScorer(pts, cols, gt, bw.PART_NAMES, part_colors=bw.PART_COLORS)
    cand = bw.candidates(base, 6)
    cand[1:, :6] = base[:6] + (cand[1:, :6] - base[:6]) * 0.05
    _, counts, scores = cpu.run(cand, processes=2)
    seg = oracle.mask_parts_from_image(gt, oracle.PART_COLORS, bw.PART_NAMES)
    sel = {p: oracle.PART_COLORS[p] for p in bw.PART_NAMES}
    for k in range(len(cand)):
        s, inter, uni = oracle.score_candidate(pts, cols, seg, sel, {"cam_pos": cand[k, 0:3], "target": cand[k, 3:6], "f": cand[k, 6],
                                                                    "cx": cand[k, 7], "cy": cand[k, 8]}, H, W)
        assert np.array_equal(counts[k, :, 0], inter) and np.array_equal(counts[k, :, 1], uni) and scores[k] == s
    if os.path.isdir("/root/reference"):
        assert cpu.kind == "reference"


def test_reference_arm_runs_without_the_product_package():
    """bench.py --impl reference on a tiny grid in a subprocess: prints one JSON line, the same `config` keys as the GPU
    arm, and never loads the product library."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid", "32", "--mask", "64",
                        "--steps", "1", "--warmup", "1", "--cpu-sample", "4"], capture_output=True, text=True, timeout=600,
                       env={**os.environ, "P3D_TRACE_IMPORTS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["blas_env"]["OPENBLAS_NUM_THREADS"] == "1"
    import bench_workload as bw
    assert line["config"] == json.loads(json.dumps(bw.bench_config(32, 64, 64, bw.PART_NAMES, line["config"]["points"], 2048)))
    assert "libp3d_b200" not in r.stderr
