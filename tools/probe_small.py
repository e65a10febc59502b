"""Segment splat vs per-point splat on the smaller configurations (run twice: with and without P3D_SPLAT_POINTS=1)."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
PKG = "part-based-3d-reconstruction_b200"
syn = importlib.import_module(PKG + ".synthetic"); ce = importlib.import_module(PKG + ".utils.camera_estimation")
cfg = importlib.import_module(PKG + ".utils.config"); mu = importlib.import_module(PKG + ".utils.mask_utils")
dev = torch.device("cuda")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rate(sc, cand, reps=3):
    cd = torch.from_numpy(cand).to(dev)
    sc.score_device(cd); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sc.score_device(cd)
    e1.record(); torch.cuda.synchronize()
    return len(cand) * reps / (e0.elapsed_time(e1) * 1e-3)


data = os.path.join(ROOT, "tests", "golden", "data")
grid = np.load(os.path.join(data, "results", "1.Orthographic_Voxel_Carving", "Taj_voxel_grid.npz"))["voxel_grid"]
cams = json.load(open(os.path.join(data, "results", "2.Perspective_Camera_Estimation", "Taj_camera_params_final.json")))
front = mu.load_mask(data, "Taj", "front", int(max(grid.shape)))
c = cams["front"]
base = np.array([*c["cam_pos"], *c["target"], c["f"], c["cx"], c["cy"]])
cand = syn.candidates(base, 4096)
gdev = torch.from_numpy(grid).to(dev)
for tag, parts in (("taj minarets", ["front_minarets", "back_minarets"]), ("taj all", syn.PART_NAMES)):
    sc = ce.CandidateScorer(gdev, front, cfg.PART_COLORS, parts)
    nseg = 0 if sc.segs is None else int(sc.segs.shape[0])
    print(tag, "points", sc.n_points, "segs", nseg, "fill", round(sc.n_points / max(8 * nseg, 1), 3), "cand/s", round(rate(sc, cand), 1))
for N, H, K in ((256, 1024, 4096), (128, 512, 4096)):
    rgb = torch.from_numpy(syn.label_lut()).to(dev)[syn.monument_labels(N, dev).long()]
    b = syn.base_camera(N, H, H)
    full = ce.CandidateScorer(rgb, torch.zeros((H, H, 3), dtype=torch.uint8, device=dev), cfg.PART_COLORS, syn.PART_NAMES)
    gt = full.render(ce.row_to_params(b + 1.0))
    sc = ce.CandidateScorer(rgb, gt, cfg.PART_COLORS, syn.PART_NAMES)
    nseg = 0 if sc.segs is None else int(sc.segs.shape[0])
    print(f"syn {N}^3 / {H}^2", "points", sc.n_points, "segs", nseg, "fill", round(sc.n_points / max(8 * nseg, 1), 3), "cand/s", round(rate(sc, syn.candidates(b, K)), 1))
