import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
PKG = "part-based-3d-reconstruction_b200"
vc = importlib.import_module(PKG + ".utils.voxel_carving_utils"); cfg = importlib.import_module(PKG + ".utils.config"); syn = importlib.import_module(PKG + ".synthetic")
nv = importlib.import_module(PKG + ".utils._native")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lab = syn.monument_labels(N, "cuda"); front = torch.flip(lab.max(dim=0).values, dims=[0]).cpu().numpy(); del lab
lut = syn.label_lut(); lut[0] = cfg.PART_COLORS["background"]; ext = lut[front]; binm = (front > 0).astype(np.uint8)
jobs = [(["full_building"], 90), (["chhatris"], 90), (["plinth"], 90), (["front_minarets"], 90), (["small_minarets"], 90), (["dome"], 90)]
g = vc.global_carve(binm, ext, 90, return_tensor=True)
import cProfile, pstats
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); p = vc.part_carve(g, ext, jobs); torch.cuda.synchronize(); print("part_carve", 1e3 * (time.perf_counter() - t0))
pr = cProfile.Profile(); pr.enable(); p = vc.part_carve(g, ext, jobs); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
