"""Public entry points run on the GPU they are told to use, whatever the current device is (the C ABI launches
on the current device's stream).  CPU half: the decorator picks the right device; GPU half (2 GPUs): carve and score on
cuda:1 while cuda:0 is current, against the same calls on cuda:0."""
import numpy as np
import pytest
import torch

from conftest import pkg


def test_on_device_decorator_resolves_the_device(monkeypatch):
    nv = pkg("utils._native")
    seen = []

    class FakeCtx:
        def __init__(self, dev):
            self.dev = dev

        def __enter__(self):
            seen.append(self.dev)

        def __exit__(self, *a):
            return False

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch.cuda, "device", FakeCtx)

    @nv.on_device
    def f(a, b=None, device=None):
        return "ran"

    assert f(1) == "ran" and seen == []                                   # nothing names a GPU: current device
    assert f(1, device="cuda:0") == "ran" and seen == []                  # already current
    assert f(1, device="cuda:1") == "ran" and seen == [torch.device("cuda:1")]
    assert f(1, device=torch.device("cuda", 3)) == "ran" and seen[-1] == torch.device("cuda:3")
    assert f(1, device="cuda") == "ran" and len(seen) == 2                 # no index: current device


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_entry_points_follow_their_device_argument(oracle):
    vc, ce, syn, cfg = pkg("utils.voxel_carving_utils"), pkg("utils.camera_estimation"), pkg("synthetic"), pkg("utils.config")
    torch.cuda.set_device(0)
    N = 64
    lab = syn.monument_labels(N)
    front = torch.flip(lab.max(dim=0).values, dims=[0]).numpy()
    lut = syn.label_lut()
    lut[0] = cfg.PART_COLORS["background"]
    ext, binm = lut[front], (front > 0).astype(np.uint8)
    g0 = vc.global_carve(binm, ext, 90, return_tensor=True, device="cuda:0")
    g1 = vc.global_carve(binm, ext, 90, return_tensor=True, device="cuda:1")
    assert g1.device == torch.device("cuda:1") and torch.equal(g0.cpu(), g1.cpu())
    jobs = [(["full_building"], 90), (["plinth"], 90), (["dome"], 90)]
    p0, p1 = vc.part_carve(g0, ext, jobs), vc.part_carve(g1, ext, jobs)     # device follows the tensor
    assert p1.device == torch.device("cuda:1") and torch.equal(p0.cpu(), p1.cpu())
    rgb = syn.label_lut()[lab.numpy()]
    base = syn.base_camera(N, 128, 128)
    gt = ce.CandidateScorer(rgb, np.zeros((128, 128, 3), np.uint8), cfg.PART_COLORS, syn.PART_NAMES, device="cuda:1").render(
        ce.row_to_params(base + 1.0))
    cand = syn.candidates(base, 40)
    s0 = ce.score_camera_candidates(rgb, gt, cfg.PART_COLORS, syn.PART_NAMES, cand, device="cuda:0")
    s1 = ce.score_camera_candidates(rgb, gt, cfg.PART_COLORS, syn.PART_NAMES, cand, device="cuda:1")
    assert np.array_equal(s0[0], s1[0]) and np.array_equal(s0[1], s1[1]) and s0[2] == s1[2]
    assert torch.cuda.current_device() == 0
