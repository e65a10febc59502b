"""Scratch: host-side profile of one global_carve call at N^3 (python tools/probe_gc.py [N])."""
import cProfile, importlib, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
PKG = "part-based-3d-reconstruction_b200"
vc = importlib.import_module(PKG + ".utils.voxel_carving_utils"); cfg = importlib.import_module(PKG + ".utils.config"); syn = importlib.import_module(PKG + ".synthetic")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lab = syn.monument_labels(N, "cuda"); front = torch.flip(lab.max(dim=0).values, dims=[0]).cpu().numpy(); del lab
lut = syn.label_lut(); lut[0] = cfg.PART_COLORS["background"]
ext = torch.from_numpy(lut[front]).cuda(); binm = (front > 0).astype(np.uint8)
for _ in range(3): vc.global_carve(binm, ext, 90, return_tensor=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): vc.global_carve(binm, ext, 90, return_tensor=True)
torch.cuda.synchronize(); print("ms/call", (time.perf_counter() - t0) / 20 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(20): vc.global_carve(binm, ext, 90, return_tensor=True)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
