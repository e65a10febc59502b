"""Timing-only experiment: the sweep with the point list re-ordered into 3-D bricks (z-buffer locality of concurrently
running CTAs).  Keys still carry the position in the permuted list, so the RESULTS of the permuted runs are not the
reference's -- this probe only asks whether the traversal order is worth building properly."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
PKG = "part-based-3d-reconstruction_b200"
syn = importlib.import_module(PKG + ".synthetic"); ce = importlib.import_module(PKG + ".utils.camera_estimation")
cfg = importlib.import_module(PKG + ".utils.config")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
HW = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
K = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
view = sys.argv[4] if len(sys.argv) > 4 else "front"
dev = torch.device("cuda:0")
H = W = HW
lut = torch.from_numpy(syn.label_lut()).to(dev)
rgb = lut[syn.monument_labels(N, dev).long()]
base = syn.base_camera(N, H, W, view)
hidden = base + np.array([3.0, -2.0, 5.0, 1.0, -1.5, 2.0, 6.0, -3.0, 2.5])
full = ce.CandidateScorer(rgb, torch.zeros((H, W, 3), dtype=torch.uint8, device=dev), cfg.PART_COLORS, syn.PART_NAMES)
gt = torch.from_numpy(full.render(ce.row_to_params(hidden))).to(dev)
del full
scorer = ce.CandidateScorer(rgb, gt, cfg.PART_COLORS, syn.PART_NAMES)
del rgb
cand = torch.from_numpy(np.ascontiguousarray(syn.candidates(base, 65536)[:K])).to(dev)
pts0, lab0 = scorer.pts.clone(), scorer.pt_label.clone()

def run(tag):
    for _ in range(2): scorer.score_device(cand)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = scorer.score_device(cand)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"N={N} mask={HW} view={view} K={K} order={tag}: {ms:.2f} ms  {K / ms * 1e3:.0f} cand/s", flush=True)

run("flat")
for brick in (16, 32, 64, 128):
    p = pts0.to(torch.int64)                       # [x, y, z]
    nb = (N + brick - 1) // brick
    bid = ((p[:, 2] // brick) * nb + (p[:, 1] // brick)) * nb + (p[:, 0] // brick)
    order = torch.sort(bid, stable=True).indices   # ascending brick id, flat index order inside a brick
    scorer.pts = pts0[order].contiguous(); scorer.pt_label = lab0[order].contiguous()
    del p, bid, order
    run(f"brick{brick}")
# columns: (y,x) blocks, all z
for blk in (32, 64):
    p = pts0.to(torch.int64)
    nb = (N + blk - 1) // blk
    bid = (p[:, 1] // blk) * nb + (p[:, 0] // blk)
    order = torch.sort(bid, stable=True).indices
    scorer.pts = pts0[order].contiguous(); scorer.pt_label = lab0[order].contiguous()
    del p, bid, order
    run(f"column{blk}")
