"""Scratch probe: f32 vs f64 candidate sweep on the 512^3 synthetic scene (python tools/probe_f32.py)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
PKG = "part-based-3d-reconstruction_b200"
syn = importlib.import_module(PKG + ".synthetic"); ce = importlib.import_module(PKG + ".utils.camera_estimation")
cfg = importlib.import_module(PKG + ".utils.config")
N, H, W, K = 512, 1024, 1024, 512
dev = torch.device("cuda")
rgb = torch.from_numpy(syn.label_lut()).to(dev)[syn.monument_labels(N, dev).long()]
base = syn.base_camera(N, H, W)
full = ce.CandidateScorer(rgb, torch.zeros((H, W, 3), dtype=torch.uint8, device=dev), cfg.PART_COLORS, syn.PART_NAMES)
gt = full.render(ce.row_to_params(base + 2.0)); del full
cand = syn.candidates(base, K)
for dt in (np.float64, np.float32):
    sc = ce.CandidateScorer(rgb, gt, cfg.PART_COLORS, syn.PART_NAMES, dtype=dt)
    cd = torch.from_numpy(cand.astype(dt)).to(dev)
    sc.score_device(cd); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sc.score_device(cd); e1.record(); torch.cuda.synchronize()
    print(np.dtype(dt).name, K / (e0.elapsed_time(e1) * 1e-3), "cand/s")
    del sc
