"""The three optimiser loops of launch_smart_aligner (camera_estimation.py:606-725) against trajectories recorded
from the LIVE reference driven through fake widgets (tests/golden/make_golden.py aligner).  The CPU test runs the
product's SmartAligner logic on top of the oracle scorer; the GPU test runs it on the CUDA scorer."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, pkg
from helpers import OracleScorer, drive_aligner


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN, "aligner_golden.npz"))


def init_params(g):
    r = g["init"]
    return {"cam_pos": r[0:3].copy(), "target": r[3:6].copy(), "f": float(r[6]), "cx": float(r[7]), "cy": float(r[8])}


PARTS = ["front_minarets", "back_minarets"]


@pytest.mark.parametrize("lock", [False, True])
def test_optimiser_logic_matches_reference_on_cpu(oracle, golden, lock, capsys):
    ce = pkg("utils.camera_estimation")
    scorer = OracleScorer(oracle, golden["grid"], golden["image"], oracle.PART_COLORS, PARTS)
    al = ce.SmartAligner(golden["grid"], golden["image"], oracle.PART_COLORS, PARTS, init_params(golden), lock_xy_equal=lock,
                         scorer=scorer)
    got = drive_aligner(al)
    tag = "lock" if lock else "free"
    for k, v in got.items():
        assert np.array_equal(v, golden[f"{tag}_{k}"]), (tag, k, v, golden[f"{tag}_{k}"])
    log = "\n".join(l for l in capsys.readouterr().out.split("\n") if "Done" in l)
    assert log == str(golden[f"{tag}_log"])


@pytest.mark.gpu
@pytest.mark.parametrize("lock", [False, True])
def test_smart_aligner_on_gpu_matches_reference(golden, lock):
    ce = pkg("utils.camera_estimation")
    cfg = pkg("utils.config")
    saved = ce.launch_smart_aligner(golden["grid"], golden["image"], cfg.PART_COLORS, parts_for_alignment=PARTS,
                                    init_params=init_params(golden), lock_xy_equal=lock)
    got = drive_aligner(saved.aligner)
    tag = "lock" if lock else "free"
    for k, v in got.items():
        assert np.array_equal(v, golden[f"{tag}_{k}"]), (tag, k)
    assert np.array_equal(np.array([*saved["cam_pos"], *saved["target"], saved["f"], saved["cx"], saved["cy"]]),
                          golden[f"{tag}_saved"])


@pytest.mark.gpu
def test_partwise_projection_iou_matches_golden(camera_golden, taj):
    ce = pkg("utils.camera_estimation")
    cfg = pkg("utils.config")
    p = taj["cams"]["front"]
    per_part, combined, counts = ce.partwise_projection_iou(taj["grid"], cfg.PART_COLORS, taj["front"], p)
    want = camera_golden["taj_front_perpart_counts"]
    assert np.array_equal(counts, want)
    assert combined == want[-1, 0] / want[-1, 1]
    assert "background" not in per_part and per_part["dome"] == want[3, 0] / want[3, 1]
