"""Scratch timing probe for the camera sweep (not the bench): python tools/probe_sweep.py N K H parts"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
PKG = "part-based-3d-reconstruction_b200"
syn = importlib.import_module(PKG + ".synthetic"); ce = importlib.import_module(PKG + ".utils.camera_estimation")
nv = importlib.import_module(PKG + ".utils._native"); eng = importlib.import_module(PKG + ".utils._engine")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
K = int(sys.argv[2]) if len(sys.argv) > 2 else 256
H = W = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
which = sys.argv[4] if len(sys.argv) > 4 else "all"
parts = syn.PART_NAMES if which == "all" else ["front_minarets", "back_minarets"]
dev = torch.device("cuda")
t0 = time.time(); labels = syn.monument_labels(N, dev); torch.cuda.synchronize(); print("gen", time.time() - t0)
lut = torch.from_numpy(syn.label_lut()).to(dev)
rgb = lut[labels.long()]; del labels
base = syn.base_camera(N, H, W)
# GT = our own render of the hidden camera (all parts)
sc_all = ce.CandidateScorer(rgb, torch.zeros((H, W, 3), dtype=torch.uint8, device=dev), syn.PART_COLORS if hasattr(syn, "PART_COLORS") else importlib.import_module(PKG + ".utils.config").PART_COLORS, syn.PART_NAMES)
hidden = base + np.array([3.0, -2.0, 5.0, 1.0, 2.0, -3.0, 4.0, 1.5, -2.5])
gt = sc_all.render({"cam_pos": hidden[0:3], "target": hidden[3:6], "f": hidden[6], "cx": hidden[7], "cy": hidden[8]})
cfg = importlib.import_module(PKG + ".utils.config")
t0 = time.time(); scorer = ce.CandidateScorer(rgb, gt, cfg.PART_COLORS, parts); torch.cuda.synchronize(); print("setup", time.time() - t0, "points", scorer.n_points)
cand = torch.from_numpy(syn.candidates(base, K)).to(dev)
for it in range(3):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
    scorer.workspace.timing(True)
    e0.record(); counts, scores, best = scorer.score_device(cand); e1.record(); torch.cuda.synchronize()
    ms, nl = scorer.workspace.timing_read(); scorer.workspace.timing(False)
    t = e0.elapsed_time(e1)
    print(f"sweep K={K} {t:.1f} ms -> {K / t * 1e3:.1f} cand/s ; splat {ms:.1f} ms in {nl} launches; point-cands/s {scorer.n_points * K / t * 1e3:.3e}; best {int(best[0])} score {float(scores[int(best[0])]):.5f}; segs {None if scorer.segs is None else int(scorer.segs.shape[0])}")
if os.environ.get("PROBE_COMPARE"):          # same candidates through the per-point splat: counts must be identical
    ref = ce.CandidateScorer(rgb, gt, cfg.PART_COLORS, parts, use_segments=False)
    c2, s2, b2 = ref.score_device(cand)
    torch.cuda.synchronize()
    for it in range(2):
        e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True); e0.record(); c2, s2, b2 = ref.score_device(cand); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1)
    print(f"per-point splat: {t:.1f} ms -> {K / t * 1e3:.1f} cand/s; counts identical: {bool(torch.equal(c2, counts))} scores identical: {bool(torch.equal(s2, scores))}")
